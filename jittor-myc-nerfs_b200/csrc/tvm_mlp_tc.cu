// Operand images of the tensor-core appearance head (tcgen05 / TMEM, sm_100a only) and its launch entry.
// The kernel itself is k_app_tc2 (tvm_app_tc2.cu); the single-group kernel of round 1 (k_app_tc, A operands in shared
// memory) is in the history (commit 15c4a98) and in profiles/r01q_notes.txt.
// Operands live in shared memory in the UMMA K-major no-swizzle ("interleaved") canonical layout: 8-row x 16-byte core
// matrices; element (r, k) of a [R x K] 16-bit operand sits at (k/8) * (R*16) + r*16 + (k%8)*2 (LBO = R*16 bytes between
// K-chunks, SBO = 128 bytes between 8-row groups).
#include "tvm_tc.cuh"

namespace tvm {

using namespace tc;

namespace tc {
__global__ void k_pack_umma_b(const float* __restrict__ w_t, int K, int K_pad, int N_real, int N, int ldw,
                              __nv_bfloat16* __restrict__ img, int split, int split_pad, int h16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K_pad * N) return;
  const int kc = i / (N * 8), rem = i % (N * 8), n = rem / 8, kk = rem % 8;
  const int k = kc * 8 + kk;
  int src = k;
  if (k >= split) src = k < split_pad ? -1 : split + (k - split_pad);
  const float v = (src >= 0 && src < K && n < N_real) ? w_t[(size_t)src * ldw + n] : 0.0f;
  reinterpret_cast<uint16_t*>(img)[i] = to16(v, h16 != 0);
}
}  // namespace tc
__global__ void k_pack_tc_f32(const TvmModel m, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128) dst[i] = m.b1[i];
  else if (i < 256) dst[i] = m.b2[i - 128];
  else if (i < 640) dst[i] = m.w3[i - 256];
  else if (i < 644) dst[i] = (i - 640) < 3 ? m.b3[i - 640] : 0.0f;
  else if (i < 692) dst[i] = m.head_bias ? m.head_bias[i - 644] : 0.0f;
}

// forward-only extension of the image (see tc::Image): b1 into row in_c of W1, b2 as K-chunks 16/17 of W2, [W3^T; b3]
__global__ void k_pack_tc_ext(const TvmModel m, int in_c, uint16_t* __restrict__ b1_img, uint16_t* __restrict__ b2x,
                              uint16_t* __restrict__ b3, int h16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128) b1_img[((in_c >> 3) * 128 + i) * 8 + (in_c & 7)] = to16(m.b1[i], h16 != 0);
  if (i < 16 * 128) {                       // element (k = 128 + 8 kc + kk, n): [(kc)][n][kk]
    const int kc = i / (128 * 8), n = (i / 8) % 128, kk = i % 8;
    b2x[i] = to16((kc == 0 && kk == 0) ? m.b2[n] : 0.0f, h16 != 0);
  }
  if (i < 144 * 16) {                       // [(k / 8)][n < 16][k % 8]
    const int kc = i / (16 * 8), n = (i / 8) % 16, k = kc * 8 + i % 8;
    float v = 0.0f;
    if (n < 3) v = k < 128 ? m.w3[n * 128 + k] : (k == 128 ? m.b3[n] : 0.0f);
    b3[i] = to16(v, h16 != 0);
  }
}

// ------------------------------------------------------------------------------------------------
static bool tc_supported(const TvmModel& m) {
  return m.n_app == 48 && m.app_dim == 27 && m.fea_pe == 2 && m.view_pe == 2 && m.feature_c == 128;
}

int launch_app_tc2(const FwdParams& P, int num_sms, cudaStream_t stream);   // tvm_app_tc2.cu

int launch_app_tc(const FwdParams& P, int num_sms, cudaStream_t stream) {
  const uint32_t mode = P.flags & TVM_MLP_MASK;
  TVM_REQUIRE(mode == TVM_MLP_BF16 || mode == TVM_MLP_FP16, "the tensor-core path takes TVM_MLP_BF16 or TVM_MLP_FP16");
  const bool h16 = mode == TVM_MLP_FP16;
  TVM_REQUIRE(tc_supported(P.m), "tensor-core appearance head supports n_app=48, app_dim=27, fea_pe=view_pe=2, "
                                 "featureC=128 (all shipped configs); use TVM_MLP_FP32 otherwise");
  TVM_REQUIRE(P.m.tc_weights != nullptr, "TvmModel.tc_weights is NULL: call tvm_pack_mlp_tc first");
  return launch_app_tc2(P, num_sms, stream);
}

}  // namespace tvm

using namespace tvm;

extern "C" size_t tvm_tc_weights_bytes(const TvmModel* m_host) {
  if (m_host == nullptr) return Image(48, 150, 32).bytes_fwd;   // probe: "is the tensor-core path built?"
  if (!tc_supported(*m_host)) return 0;
  return Image(m_host->n_app, in_mlp_c(*m_host), head_ld(*m_host)).bytes_fwd;
}

extern "C" int tvm_pack_mlp_tc(const TvmModel* m_host, void* out, uint32_t flags, void* stream_) {
  const uint32_t mode = flags & TVM_MLP_MASK;
  TVM_REQUIRE(mode == TVM_MLP_BF16 || mode == TVM_MLP_FP16, "tvm_pack_mlp_tc: flags must name TVM_MLP_BF16 or TVM_MLP_FP16");
  const int h16 = mode == TVM_MLP_FP16;
  TVM_REQUIRE(m_host && out, "bad arguments");
  if (int rc = validate_model(*m_host)) return rc;
  TVM_REQUIRE(tc_supported(*m_host), "unsupported shape for the tensor-core appearance head");
  cudaStream_t s = (cudaStream_t)stream_;
  const int in_c = in_mlp_c(*m_host);
  const int nh = head_ld(*m_host);
  const Image img(m_host->n_app, in_c, nh);
  uint8_t* o = (uint8_t*)out;
  auto launch = [&](const float* w_t, int K, int K_pad, int N_real, int N, int ldw, uint32_t off) {
    const int n = K_pad * N;
    k_pack_umma_b<<<(n + 255) / 256, 256, 0, s>>>(w_t, K, K_pad, N_real, N, ldw, (__nv_bfloat16*)(o + off), K_pad, K_pad, h16);
  };
  launch(m_host->basis_t, img.K0, img.K0, nh, nh, nh, img.off_b0);
  launch(m_host->w1_t, in_c, img.K1, 128, 128, kFeatureC, img.off_b1);
  launch(m_host->w2_t, 128, 128, 128, 128, kFeatureC, img.off_b2);
  k_pack_tc_f32<<<3, 256, 0, s>>>(*m_host, (float*)(o + img.off_f32));
  k_pack_tc_ext<<<(144 * 16 + 255) / 256, 256, 0, s>>>(*m_host, in_c, (uint16_t*)(o + img.off_b1), (uint16_t*)(o + img.off_b2x),
                                                    (uint16_t*)(o + img.off_b3), h16);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
