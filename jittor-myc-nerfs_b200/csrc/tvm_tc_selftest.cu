// Known-answer self-test of the tcgen05 operand conventions the tensor-core kernels rely on (sm_100a only).
// Checks, against a host matmul, the two ways one shared-memory image in the K-major no-swizzle core-matrix layout
//   element (r, c) of [R x C] bf16 at (c/8) * (R*16) + r*16 + (c%8)*2
// is consumed:
//   K-major  (forward):   D[r][n]  = sum_c  A[r][c] * B[n][c]        LBO = R*16 (next 8 columns), SBO = 128 (next 8 rows)
//   MN-major (backward):  D[m][n]  = sum_r  P[r][m] * Q[r][n]        LBO = 128 (next 8 reduction rows), SBO = R*16 (next 8 columns)
//                         i.e. the SAME bytes read "transposed": activations as weight-gradient operands (reduction over
//                         the tile rows) and forward weight images as data-gradient operands (no second copy).
#include "tvm_tc.cuh"

namespace tvm {
namespace tc {

constexpr int kStRows = 128;

__device__ __forceinline__ void fill_image(uint8_t* img, const float* __restrict__ src, int R, int C) {
  for (int i = threadIdx.x; i < R * C; i += blockDim.x) {
    const int r = i / C, c = i % C;
    *reinterpret_cast<__nv_bfloat16*>(img + (c / 8) * (R * 16) + r * 16 + (c % 8) * 2) = __float2bfloat16_rn(src[i]);
  }
}

// P [128 x 128], Q [128 x 160], W [128 x 128] (fp32, row-major, values exactly representable in bf16)
//   D1 [128 x 160] = P^T Q             (both operands MN-major, reduction over the 128 rows)
//   D2 [128 x 128] = P W                (A K-major; B = W read MN-major: D2[r][j] = sum_o P[r][o] W[o][j])
//   D3 [128 x 128] = P W^T              (both K-major: the forward convention, D3[r][o] = sum_c P[r][c] W[o][c])
__global__ void __launch_bounds__(128, 1) k_selftest_umma(const float* __restrict__ P, const float* __restrict__ Q,
                                                          const float* __restrict__ W, float* __restrict__ D1,
                                                          float* __restrict__ D2, float* __restrict__ D3) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sP = smem;                           // [128 x 128] bf16, 32 KB
  uint8_t* sQ = sP + 128 * 128 * 2;             // [128 x 160] bf16, 40 KB
  uint8_t* sW = sQ + 128 * 160 * 2;             // [128 x 128] bf16, 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sW + 128 * 128 * 2);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  fill_image(sP, P, 128, 128);
  fill_image(sQ, Q, 128, 160);
  fill_image(sW, W, 128, 128);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t R16 = kStRows * 16;
  if (tid == 0) {
    // D1: 8 reduction steps of 16 rows (256 bytes apart)
    for (int s = 0; s < 8; ++s)
      umma_bf16(tmem, smem_desc(smem_u32(sP) + s * 256, 128, R16), smem_desc(smem_u32(sQ) + s * 256, 128, R16),
                instr_desc(128, 160) | kIdescAMajorMN | kIdescBMajorMN, s > 0);
    // D2: A = P K-major (steps of 16 columns), B = W image read MN-major (reduction over its rows o)
    for (int s = 0; s < 8; ++s)
      umma_bf16(tmem + 160, smem_desc(smem_u32(sP) + s * 2 * R16, R16, 128), smem_desc(smem_u32(sW) + s * 256, 128, R16),
                instr_desc(128, 128) | kIdescBMajorMN, s > 0);
    // D3: forward convention
    for (int s = 0; s < 8; ++s)
      umma_bf16(tmem + 288, smem_desc(smem_u32(sP) + s * 2 * R16, R16, 128), smem_desc(smem_u32(sW) + s * 2 * R16, R16, 128),
                instr_desc(128, 128), s > 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  fence_after();
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  float v[32];
  for (int cb = 0; cb < 5; ++cb) {
    tmem_ld32(lane_addr + cb * 32, v);
    for (int i = 0; i < 32; ++i) D1[tid * 160 + cb * 32 + i] = v[i];
  }
  for (int cb = 0; cb < 4; ++cb) {
    tmem_ld32(lane_addr + 160 + cb * 32, v);
    for (int i = 0; i < 32; ++i) D2[tid * 128 + cb * 32 + i] = v[i];
  }
  for (int cb = 0; cb < 4; ++cb) {
    tmem_ld32(lane_addr + 288 + cb * 32, v);
    for (int i = 0; i < 32; ++i) D3[tid * 128 + cb * 32 + i] = v[i];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace tvm

using namespace tvm;

extern "C" int tvm_selftest_umma(const float* P, const float* Q, const float* W, float* D1, float* D2, float* D3, void* stream) {
  TVM_REQUIRE(P && Q && W && D1 && D2 && D3, "null argument");
  const size_t smem = (128 * 128 + 128 * 160 + 128 * 128) * 2 + 64;
  TVM_CHECK_CUDA(cudaFuncSetAttribute(tc::k_selftest_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::k_selftest_umma<<<1, 128, smem, (cudaStream_t)stream>>>(P, Q, W, D1, D2, D3);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
