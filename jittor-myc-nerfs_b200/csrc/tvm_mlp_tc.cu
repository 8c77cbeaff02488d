// tcgen05 tensor-core appearance head (placeholder until the UMMA path lands).
#include "tvm_common.cuh"
namespace tvm {
int launch_app_tc(const FwdParams& P, int num_sms, cudaStream_t stream) {
  (void)P; (void)num_sms; (void)stream;
  set_error("tensor-core appearance head not built in this library");
  return -3;
}
}  // namespace tvm
extern "C" size_t tvm_tc_weights_bytes(const TvmModel* m_host) { (void)m_host; return 0; }
extern "C" int tvm_pack_mlp_tc(const TvmModel* m_host, void* out, void* stream) {
  (void)m_host; (void)out; (void)stream;
  tvm::set_error("tensor-core appearance head not built in this library");
  return -3;
}
