// Backward pass (placeholder).
#include "tvm_common.cuh"
extern "C" int tvm_backward(const TvmModel* m_host, const float* rays, int n_rays, int n_samples,
                 const float* jitter, uint32_t flags, const float* rgb_map, const float* d_rgb_map,
                 const TvmGrads* grads_host, void* ws, size_t ws_bytes, void* stream) {
  tvm::set_error("backward not built in this library");
  return -3;
}
