// tcgen05 / TMEM / mbarrier PTX wrappers shared by the tensor-core kernels (sm_100a only).
// Operands live in shared memory in the UMMA K-major no-swizzle ("interleaved") canonical layout:
// 8-row x 16-byte core matrices; element (r, k) of a [R x K] bf16 operand sits at
//   (k/8) * (R*16) + r*16 + (k%8)*2        (LBO = R*16 bytes between K-chunks, SBO = 128 bytes between 8-row groups)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tvm_common.cuh"

namespace tvm {
namespace tc {

constexpr int kRows = 128;          // UMMA M

// Rows per tile of a persistent launch over n_ent entries whose CTAs take `slots` tiles per round (grid x tile pipelines per
// CTA).  128 -- the MMA's M -- when the launch is many rounds long; for the few-round launches of a training step (4096 rays:
// ~240 tiles on 148 SMs = 1.6 rounds) the multiple of 8 that spreads the entries evenly over WHOLE rounds, so that no SM
// idles through a second round while the others work on full tiles (rows behind it are dead: zero operands, no gather, no
// scatter; the MMAs run on all 128 rows either way).
__device__ __forceinline__ uint32_t balanced_tile_rows(uint32_t n_ent, uint32_t slots) {
#ifdef TVM_NO_BALANCED_TILES
  return kRows;
#endif
  const uint32_t t = (n_ent + kRows - 1) / kRows;
  if (t == 0 || t >= 8u * slots) return kRows;
  const uint32_t t2 = (t + slots - 1) / slots * slots;
  const uint32_t r = ((n_ent + t2 - 1) / t2 + 7u) & ~7u;
  return r < (uint32_t)kRows ? r : (uint32_t)kRows;
}

constexpr int kTmemCols = 256;      // [0,128): layer accumulators, [128,160): basis accumulator
constexpr int kColBasis = 128;

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  // try_wait suspends the thread in hardware until the phase completes or the time hint (ns) runs out, so a waiting
  // warp costs one issue slot per ~20 us instead of one spin iteration per ~50 cycles (the spin loops of the waiting MLP
  // threads were 12-19 % of all instructions k_app_tc2 executed before the hint, profiles/r02_notes.txt)
#ifndef TVM_MBAR_HINT_NS
#define TVM_MBAR_HINT_NS 20000
#endif
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity), "r"((uint32_t)TVM_MBAR_HINT_NS)
        : "memory");
    if (spin > (1u << 20)) __trap();   // a lost tcgen05.commit must fail the launch, not hang the GPU
  }
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (bf16 inputs, fp32 accumulate), M=128, K=16
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t gets lane (base lane + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- 16-bit A operands kept in TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc) ------------------------------------
// Layout of a [128 x K] 16-bit A operand in TMEM: row r = lane r; 32-bit column c holds K elements (2c, 2c+1), even k in the
// low half; one K=16 step of the MMA consumes 8 columns.  A thread (= row) writes its packed activations with tcgen05.st
// and never touches shared memory: no STS wavefronts and no A-operand fetches on the L1 data pipe for that layer.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16, M = 128, K = 16
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// K-major, no swizzle: LBO = byte distance between K-adjacent core matrices, SBO = between 8-row groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// bf16 x bf16 -> f32, both operands K-major
// h16: operands are IEEE fp16 (format 0) instead of bf16 (format 1)
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, bool h16 = false) {
  return (1u << 4) | ((h16 ? 0u : 1u) << 7) | ((h16 ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// operand read "transposed" (MN-major): the reduction index runs over the image ROWS; LBO = 128, SBO = rows * 16
constexpr uint32_t kIdescAMajorMN = 1u << 15;
constexpr uint32_t kIdescBMajorMN = 1u << 16;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 16-bit operand element pairs of the appearance head: bf16 (TVM_MLP_BF16) or fp16 (TVM_MLP_FP16)
template <bool H16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if (H16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16(a, b);
}
// the same pack with ReLU fused into the conversion (cvt.rn.relu.*x2.f32); `lo` lands in the low half
template <bool H16>
__device__ __forceinline__ uint32_t pack16_relu(float lo, float hi) {
  uint32_t d;
  if (H16) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool H16>
__device__ __forceinline__ float2 unpack16(uint32_t v) {
  if (H16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
// Mixed-precision multiply-add of sm_100: d (fp32) = a (16 bit) * b (16 bit) + c (fp32), one instruction (SASS FHFMA), each
// 16-bit operand either half of a 32-bit register -- a packed pair is consumed without an unpack.
template <bool H16>
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float c) {
  float d;
  if (H16) asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  else asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint16_t half_of(uint32_t v, bool hi) { return hi ? (uint16_t)(v >> 16) : (uint16_t)(v & 0xffffu); }
// a value that is exact in the 16-bit format (lattice weights), as its bit pattern
template <bool H16>
__device__ __forceinline__ uint16_t w16(float v) {
  if (H16) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__host__ __device__ inline uint16_t to16(float v, bool h16) {
  if (h16) {
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const uint16_t*>(&h);
  }
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t*>(&b);
}


// Byte layout of the weight image built by tvm_pack_mlp_tc (copied verbatim into shared memory)
struct Image {
  int K0, K1, NH;                // padded reduction lengths of GEMM0 / GEMM1 (multiples of 16); GEMM0 width
  uint32_t off_b0, off_b1, off_b2, off_f32, bytes;
  // fp32 tail: b1[128] b2[128] w3[3][128] b3[4] head_bias[48]
  // Forward-only extension behind `bytes` (the backward kernel copies [0, bytes) only):
  //   off_b2x  K-chunks 16, 17 of the W2 operand: row 128 = b2 (the forward feeds a constant-one column there)
  //   off_b3   W3 operand [144][16]: rows 0..127 = mlp.4.weight^T (3 real columns), row 128 = b3
  // and row in_c of the W1 operand (inside its K padding) holds b1: the biases and the 128 -> 3 layer ride in the GEMMs.
  uint32_t off_b2x, off_b3, bytes_fwd;
  __host__ __device__ Image(int n_app, int in_c, int nh) {
    K0 = 3 * n_app;
    K1 = (in_c + 15) / 16 * 16;
    NH = nh;
    off_b0 = 0;
    off_b1 = off_b0 + (uint32_t)K0 * NH * 2;
    off_b2 = off_b1 + (uint32_t)K1 * 128 * 2;
    off_f32 = off_b2 + 128u * 128 * 2;
    bytes = off_f32 + (128 + 128 + 3 * 128 + 4 + 48) * 4;
    off_b2x = (bytes + 15u) & ~15u;
    off_b3 = off_b2x + 16u * 128 * 2;
    bytes_fwd = off_b3 + 144u * 16 * 2;
  }
};
constexpr int kColOut = 192;        // TMEM columns [192, 208): accumulator of the 128 -> 3 layer (forward kernel)

// fp32 [K][ldw] (row j = input j) -> bf16 UMMA image [(K_pad/8)][N][8]; out-of-range entries are 0.
// Image column k reads source row k for k < split, nothing for split <= k < split_pad, and row
// split + (k - split_pad) beyond (used to pad the first block of a concatenated input to 16).
__global__ void k_pack_umma_b(const float* __restrict__ w_t, int K, int K_pad, int N_real, int N, int ldw,
                              __nv_bfloat16* __restrict__ img, int split, int split_pad, int h16 = 0);

}  // namespace tc
}  // namespace tvm
