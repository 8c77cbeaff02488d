// NeRF++ background MLP (models/nerfplusplus.py:66-140, :285-310) on the 5th-generation tensor cores
// (tcgen05 / TMEM), sm_100a only.  Replaces k_bg_simt when TVM_MLP_BF16 is requested.
//
// One persistent CTA per SM holds the whole network (five bf16 B operands, 106 KB) in shared memory and
// runs THREE independent 128-thread pipelines (thread = sample = TMEM lane), so that the MMAs of one
// pipeline overlap the epilogues of the others (3 x 40 KB A operands: the 227 KB of shared memory are full).  A pipeline takes one active ray at a time (dynamic
// scheduling through a global counter) and walks its 512 samples front to back in 4 tiles of 128:
//
//   geometry  depth2pts_outside + Embedder (:207-237, :40-56)  -> A[:, 0:32)   = [emb 20 | 1.0 | 0...]  bf16
//   L0   A[:, 0:32)            . W0p   (row 20 = b0)            -> TMEM, ReLU  -> A[:, 32:160)
//   L1   A[:, 32:160) + A[:, 16:32) . W1p (bias row)            -> TMEM, ReLU  -> A[:, 32:160)
//   L2   A[:, 0:160)           . W2p   (skip connection: [input_pts | base], row 20 = b2) -> ReLU -> A[:, 32:160)
//   L3   A[:, 32:160)          . [Wf | w_sigma]  (N = 80; base_remap folded into rgb_layers.0 by tvm_bg_fold)
//          epilogue: + per-ray view bias, ReLU -> A[:, 32:96);  sigma = |acc[64] + b_sigma|
//   L4   A[:, 32:96)           . W_rgb (N = 16)                 -> sigmoid
//   composite  alpha = 1 - exp(-sigma dz), warp prefix products + cross-warp carry, early termination at T < 1e-6
//
// Biases ride in the GEMMs through the constant-one column 20 of the position block (free K padding).
// ReLU is fused into the fp32 -> bf16 pack (cvt.rn.relu.bf16x2.f32).
#include "tvm_bg.cuh"
#include "tvm_tc.cuh"

namespace tvm {
namespace bgtc {

using namespace tc;

constexpr int kPipes = 3;
constexpr int kPipeThreads = 128;
constexpr int kThreads = kPipes * kPipeThreads;
using namespace bgimg;                  // operand shapes and byte offsets of the weight image (tvm_bg.cuh)
constexpr int kTmemColsBg = 512;        // 128 accumulator columns per pipeline (power of two >= 3 * 128)

constexpr uint32_t kLboA = kRows * 16;
constexpr uint32_t kABytes = kRows * kAK * 2;

__device__ __forceinline__ void pipe_sync(int pipe) {
  asm volatile("bar.sync %0, %1;" ::"r"(pipe + 1), "n"(kPipeThreads) : "memory");
}
// two fp32 -> packed bf16 with ReLU; `lo` lands in the low half (= first in memory)
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// accumulator columns [0,128) of this thread's lane -> ReLU -> bf16 -> A[:, 32:160)
__device__ __forceinline__ void epi_relu_store(uint32_t lane_addr, uint8_t* arow_hidden) {
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float y[32];
    tmem_ld32(lane_addr + cb * 32, y);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 v;
      v.x = pack_relu_bf16(y[g * 8 + 0], y[g * 8 + 1]);
      v.y = pack_relu_bf16(y[g * 8 + 2], y[g * 8 + 3]);
      v.z = pack_relu_bf16(y[g * 8 + 4], y[g * 8 + 5]);
      v.w = pack_relu_bf16(y[g * 8 + 6], y[g * 8 + 7]);
      *reinterpret_cast<uint4*>(arow_hidden + (cb * 4 + g) * kLboA) = v;
    }
  }
}

// issue `steps` K-steps (16 columns each) of D (+)= A . B; A step s starts at a_base + s * 2 * LBO_A
__device__ __forceinline__ void issue_steps(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t lbo_b, int steps,
                                            uint32_t idesc, bool accumulate_first) {
#pragma unroll
  for (int s = 0; s < steps; ++s)
    umma_bf16(tmem_d, smem_desc(a_base + s * 2 * kLboA, kLboA, 128), smem_desc(b_base + s * 2 * lbo_b, lbo_b, 128), idesc,
              (accumulate_first || s > 0) ? 1u : 0u);
}

__global__ void __launch_bounds__(kThreads, 1) k_bg_tc(const FwdParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + kImageBytes;                                  // kPipes x [128 x 160] bf16
  float* sVB = reinterpret_cast<float*>(sA + kPipes * kABytes);      // [kPipes][64] per-ray bias of the hidden rgb layer
  float* sCS = sVB;                                                  // per-warp colour sums alias VB[0..15] (VB is dead by then)
  float* sWP = sVB + kPipes * kBgHid;                                // [kPipes][4] per-warp transmittance products
  uint64_t* bars = reinterpret_cast<uint64_t*>(sWP + kPipes * 4);    // [kPipes] MMA completion
  uint32_t* sRay = reinterpret_cast<uint32_t*>(bars + kPipes);       // [kPipes] ray index handed out by the scheduler
  uint32_t* tmem_slot = sRay + kPipes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pipe = tid / kPipeThreads, ptid = tid % kPipeThreads, pwarp = ptid >> 5;
  const TvmBgNet& bg = P.bg;
  const float R = P.m.radii;

  {
    const uint4* src = reinterpret_cast<const uint4*>(bg.tc_weights);
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (uint32_t i = tid; i < kImageBytes / 16; i += kThreads) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    for (int p = 0; p < kPipes; ++p) mbar_init(&bars[p], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemColsBg);
  // columns [21, 32) of the position block stay zero for the whole kernel; column 20 is the constant one
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot + (uint32_t)pipe * 128u;                    // this pipeline's accumulator columns
  const uint32_t lane_addr = tmem + ((uint32_t)(pwarp * 32) << 16);
  uint8_t* A = sA + pipe * kABytes;
  uint8_t* arow = A + ptid * 16;
  uint8_t* arow_hidden = arow + (kPosK / 8) * kLboA;
  const uint32_t aA = smem_u32(A), aAh = aA + (kPosK / 8) * kLboA;
  const uint32_t aW0 = smem_u32(sW + kOffW0), aW1 = smem_u32(sW + kOffW1), aW2 = smem_u32(sW + kOffW2),
                 aW3 = smem_u32(sW + kOffW3), aW4 = smem_u32(sW + kOffW4);
  const float b_sigma = reinterpret_cast<const float*>(sW + kOffF32)[0];
  const float b_r = reinterpret_cast<const float*>(sW + kOffF32)[1], b_g = reinterpret_cast<const float*>(sW + kOffF32)[2],
              b_b = reinterpret_cast<const float*>(sW + kOffF32)[3];
  constexpr uint32_t ID128 = instr_desc(128, 128), ID80 = instr_desc(128, kN3), ID16 = instr_desc(128, kN4);
  uint64_t* mma_bar = &bars[pipe];
  uint32_t phase = 0;
  float* VB = sVB + pipe * kBgHid;
  float* WP = sWP + pipe * 4;
  float* CS = sCS + pipe * kBgHid;
  const uint32_t n_active = P.ws.n_entries[1];

  auto mma_wait = [&]() {
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    fence_after();
  };
  auto publish_A = [&]() {      // generic-proxy writes of A -> visible to the tensor core, then pipeline barrier
    fence_async_smem();
    fence_before();
    pipe_sync(pipe);
  };

  while (true) {
    if (ptid == 0) sRay[pipe] = atomicAdd(P.ws.n_entries + 2, 1u);
    pipe_sync(pipe);
    const uint32_t idx = sRay[pipe];
    if (idx >= n_active) break;
    const uint32_t ray = P.ws.bg_list[idx];
    const float* ray6 = P.rays + 6 * (size_t)ray;
    const float* rnd = P.bg_rand + (size_t)ray * kBgSamples;
    BgRay g;
    bg_ray_setup(ray6, R, g);
    if (ptid < kBgHid) {
      // view-direction embedding through its slice of rgb_layers.0, plus the folded bias (fp32, once per ray)
      const float dn = 1.0f / sqrtf(ray6[3] * ray6[3] + ray6[4] * ray6[4] + ray6[5] * ray6[5]);
      const float v[3] = {ray6[3] * dn, ray6[4] * dn, ray6[5] * dn};
      float e[kDirDim];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        e[c] = v[c];
        e[3 + c] = sinf(v[c]);
        e[6 + c] = cosf(v[c]);
        e[9 + c] = sinf(2.0f * v[c]);
        e[12 + c] = cosf(2.0f * v[c]);
      }
      float a = bg.bf[ptid];
#pragma unroll
      for (int j = 0; j < kDirDim; ++j) a = fmaf(e[j], __ldg(bg.wv_t + j * kBgHid + ptid), a);
      VB[ptid] = a;
    }
    float T = 1.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
    int tiles_done = 0;
    // geometry + embedding of one sample: flipped order j (nearest the sphere first), original index i = 511 - j
    float nxt_dz;
    uint4 nxt_e0, nxt_e1, nxt_e2;
    auto geometry = [&](int tile) {
      const int j = tile * kRows + ptid, i = kBgSamples - 1 - j;
      const float z = bg_depth(i, R, rnd);
      nxt_dz = (i > 0) ? z - bg_depth(i - 1, R, rnd) : 1e10f;             // bg_dists, HUGE_NUMBER last (:299-300)
      const float theta = asinf(g.pmn * z / (R * R));
      float sa, ca;
      __sincosf(g.phi - theta, &sa, &ca);
      float x[4], s1[4], cc1[4], s2[4], cc2[4];
#pragma unroll
      for (int c = 0; c < 3; ++c) x[c] = g.p_sphere[c] * ca + g.cross_ap[c] * sa + g.axis[c] * g.axis_dot * (1.0f - ca);
      x[3] = z;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        __sincosf(x[c], &s1[c], &cc1[c]);
        s2[c] = 2.0f * s1[c] * cc1[c];
        cc2[c] = 1.0f - 2.0f * s1[c] * s1[c];
      }
      nxt_e0 = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(s1[0], s1[1]), pack_bf16(s1[2], s1[3]));
      nxt_e1 = make_uint4(pack_bf16(cc1[0], cc1[1]), pack_bf16(cc1[2], cc1[3]), pack_bf16(s2[0], s2[1]), pack_bf16(s2[2], s2[3]));
      nxt_e2 = make_uint4(pack_bf16(cc2[0], cc2[1]), pack_bf16(cc2[2], cc2[3]), pack_bf16(1.0f, 0.0f), 0u);
    };
    geometry(0);

#pragma unroll 1
    for (int tile = 0; tile < kBgSamples / kRows; ++tile) {
      ++tiles_done;
      // ---- position block of this tile (prefetched during the previous tile's L4) ------------------------------------
      const float dz = nxt_dz;
      *reinterpret_cast<uint4*>(arow + 0 * kLboA) = nxt_e0;
      *reinterpret_cast<uint4*>(arow + 1 * kLboA) = nxt_e1;
      *reinterpret_cast<uint4*>(arow + 2 * kLboA) = nxt_e2;
      *reinterpret_cast<uint4*>(arow + 3 * kLboA) = make_uint4(0u, 0u, 0u, 0u);
      publish_A();
      // ---- L0 ------------------------------------------------------------------------------------------------
      if (ptid == 0) {
        fence_after();
        issue_steps(tmem, aA, aW0, kFeatureC * 16, kPosK / 16, ID128, false);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_relu_store(lane_addr, arow_hidden);
      publish_A();
      // ---- L1: hidden (8 steps) + position columns 16..31 (bias row) -----------------------------------------------
      if (ptid == 0) {
        fence_after();
        issue_steps(tmem, aAh, aW1, kFeatureC * 16, kFeatureC / 16, ID128, false);
        issue_steps(tmem, aA + 2 * kLboA, aW1 + (kFeatureC / 8) * kFeatureC * 16, kFeatureC * 16, 1, ID128, true);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_relu_store(lane_addr, arow_hidden);
      publish_A();
      // ---- L2: [input_pts | base] (skip connection), bias in row 20 ----------------------------------------------------
      if (ptid == 0) {
        fence_after();
        issue_steps(tmem, aA, aW2, kFeatureC * 16, kAK / 16, ID128, false);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_relu_store(lane_addr, arow_hidden);
      publish_A();
      // ---- L3: hidden rgb layer (base_remap folded in) and sigma ---------------------------------------------------------
      if (ptid == 0) {
        fence_after();
        issue_steps(tmem, aAh, aW3, kN3 * 16, kFeatureC / 16, ID80, false);
        umma_commit(mma_bar);
      }
      mma_wait();
      float sigma;
      {
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          float y[32];
          tmem_ld32(lane_addr + cb * 32, y);
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const float4 b = *reinterpret_cast<const float4*>(VB + cb * 32 + q);
            y[q] += b.x; y[q + 1] += b.y; y[q + 2] += b.z; y[q + 3] += b.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = pack_relu_bf16(y[q * 8 + 0], y[q * 8 + 1]);
            v.y = pack_relu_bf16(y[q * 8 + 2], y[q * 8 + 3]);
            v.z = pack_relu_bf16(y[q * 8 + 4], y[q * 8 + 5]);
            v.w = pack_relu_bf16(y[q * 8 + 6], y[q * 8 + 7]);
            *reinterpret_cast<uint4*>(arow_hidden + (cb * 4 + q) * kLboA) = v;
          }
        }
        float t16[16];
        tmem_ld16(lane_addr + 64, t16);
        sigma = fabsf(t16[0] + b_sigma);                                     // sigma = |w . base + b| (:128-129)
      }
      publish_A();
      // ---- L4: 64 -> 3 ---------------------------------------------------------------------------------------------------
      if (ptid == 0) {
        fence_after();
        issue_steps(tmem, aAh, aW4, kN4 * 16, kBgHid / 16, ID16, false);
        umma_commit(mma_bar);
      }
      if (tile + 1 < kBgSamples / kRows) geometry(tile + 1);       // overlaps the L4 round trip
      mma_wait();
      float rgb[3];
      {
        float t16[16];
        tmem_ld16(lane_addr, t16);
        rgb[0] = 1.0f / (1.0f + __expf(-(t16[0] + b_r)));
        rgb[1] = 1.0f / (1.0f + __expf(-(t16[1] + b_g)));
        rgb[2] = 1.0f / (1.0f + __expf(-(t16[2] + b_b)));
      }
      // ---- front-to-back compositing of the tile (:301-310) -------------------------------------------------------------------
      const float al = 1.0f - expf(-sigma * dz);
      const float vv = 1.0f - al + 1e-6f;                                   // TINY_NUMBER (:303)
      float pref = vv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, pref, o);
        if (lane >= o) pref *= t;
      }
      float excl = __shfl_up_sync(0xffffffffu, pref, 1);
      if (lane == 0) excl = 1.0f;
      if (lane == 31) WP[pwarp] = pref;
      fence_before();
      pipe_sync(pipe);                 // also: every lane has drained its TMEM loads before the next L0 is issued
      const float p0 = WP[0], p1 = WP[1], p2 = WP[2], p3 = WP[3];
      const float before = pwarp == 0 ? 1.0f : pwarp == 1 ? p0 : pwarp == 2 ? p0 * p1 : p0 * p1 * p2;
      const float w = al * (T * before * excl);
      c0 = fmaf(w, rgb[0], c0);
      c1 = fmaf(w, rgb[1], c1);
      c2 = fmaf(w, rgb[2], c2);
      T = T * (p0 * p1 * p2 * p3);
      if (T < 1e-6f) break;            // uniform over the pipeline: remaining weight mass < 1e-6
      // WP is rewritten only after the next tile's five barriers, so no extra barrier is needed here
    }
    c0 = warp_sum(c0);
    c1 = warp_sum(c1);
    c2 = warp_sum(c2);
    if (lane == 0) {
      CS[pwarp * 4 + 0] = c0;
      CS[pwarp * 4 + 1] = c1;
      CS[pwarp * 4 + 2] = c2;
    }
    pipe_sync(pipe);
    if (ptid == 3 && P.counters) {
      atomicAdd(&P.counters[TVM_CNT_BG_RAYS], 1ull);
      atomicAdd(&P.counters[TVM_CNT_BG_SAMPLES], (unsigned long long)(tiles_done * kRows));
    }
    if (ptid < 3) {
      const float c = (CS[ptid] + CS[4 + ptid]) + (CS[8 + ptid] + CS[12 + ptid]);
      const float lam = P.ws.bg_lambda[ray];
      P.rgb_map[(size_t)ray * 3 + ptid] += lam * c;                       // nerfplusplus.py:314-317 (no clamp after the sum)
      P.ws.bg_rgb[(size_t)ray * 3 + ptid] = c;                            // kept for tvm_backward_npp
      if (P.aux.bg_rgb_map) P.aux.bg_rgb_map[(size_t)ray * 3 + ptid] = c;
    }
    // the scheduler barrier at the top of the loop orders these CS/VB reads before the next ray's writes
  }

  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, kTmemColsBg);
}

// ---- weight image --------------------------------------------------------------------------------------------------
// value of B operand `layer` at (reduction index k, output n); the K order of every layer is the order in which
// k_bg_tc issues its K-steps
__device__ float bg_weight(const TvmBgNet& b, int layer, int k, int n) {
  switch (layer) {
    case 0:   // K = 32: position block
      if (k < kPosDim) return b.w0_t[k * kFeatureC + n];
      return k == kOneCol ? b.b0[n] : 0.0f;
    case 1:   // K = 144: hidden 0..127, then position columns 16..31 (column 20 -> k = 132)
      if (k < kFeatureC) return b.w1_t[k * kFeatureC + n];
      return k == kFeatureC + (kOneCol - 16) ? b.b1[n] : 0.0f;
    case 2:   // K = 160: position block (input_pts rows of base_layers.2), then hidden
      if (k < kPosDim) return b.w2_t[k * kFeatureC + n];
      if (k < kPosK) return k == kOneCol ? b.b2[n] : 0.0f;
      return b.w2_t[(kPosDim + k - kPosK) * kFeatureC + n];
    case 3:   // K = 128, N = 80: [wf | w_sigma | 0]
      if (n < kBgHid) return b.wf_t[k * kBgHid + n];
      return n == kBgHid ? b.w_sigma[k] : 0.0f;
    default:  // K = 64, N = 16: rgb_layers.2.weight^T
      return n < 3 ? b.w_rgb[n * kBgHid + k] : 0.0f;
  }
}

__global__ void k_pack_bg_tc(const TvmBgNet b, uint8_t* __restrict__ out) {
  const int layer = blockIdx.y;
  const int K = layer == 0 ? kPosK : layer == 1 ? kK1 : layer == 2 ? kAK : layer == 3 ? kFeatureC : kBgHid;
  const int N = layer < 3 ? kFeatureC : layer == 3 ? kN3 : kN4;
  const uint32_t off = layer == 0 ? kOffW0 : layer == 1 ? kOffW1 : layer == 2 ? kOffW2 : layer == 3 ? kOffW3 : kOffW4;
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(out + off);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * N; i += gridDim.x * blockDim.x) {
    const int kc = i / (N * 8), rem = i % (N * 8), n = rem / 8, kk = rem % 8;
    img[i] = __float2bfloat16_rn(bg_weight(b, layer, kc * 8 + kk, n));
  }
  if (layer == 0 && blockIdx.x == 0 && threadIdx.x < 4) {
    float* f = reinterpret_cast<float*>(out + kOffF32);
    f[threadIdx.x] = threadIdx.x == 0 ? b.b_sigma[0] : b.b_rgb[threadIdx.x - 1];
  }
}

}  // namespace bgtc

int launch_bg_tc(const FwdParams& P, int num_sms, cudaStream_t stream) {
  using namespace bgtc;
  TVM_REQUIRE(P.bg.tc_weights != nullptr, "TvmBgNet.tc_weights is NULL: call tvm_pack_bg_tc first");
  const size_t smem = kImageBytes + kPipes * kABytes + kPipes * (kBgHid + 4) * 4 + kPipes * 8 + kPipes * 4 + 16;
  static_assert(kImageBytes % 16 == 0 && (kPipes * 4 * 4) % 8 == 0, "shared-memory carve-up alignment");
  TVM_REQUIRE(smem <= 227 * 1024, "k_bg_tc shared memory");
  TVM_CHECK_CUDA(cudaFuncSetAttribute(k_bg_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bg_tc<<<num_sms, kThreads, smem, stream>>>(P);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvm

using namespace tvm;

extern "C" size_t tvm_bg_tc_bytes(void) { return bgtc::kImageBytes; }

extern "C" int tvm_pack_bg_tc(const TvmBgNet* bg_host, void* out, void* stream) {
  TVM_REQUIRE(bg_host && out, "bad arguments");
  const TvmBgNet& b = *bg_host;
  TVM_REQUIRE(b.w0_t && b.b0 && b.w1_t && b.b1 && b.w2_t && b.b2 && b.w_sigma && b.b_sigma && b.wf_t && b.bf && b.wv_t &&
              b.w_rgb && b.b_rgb, "null TvmBgNet pointer");
  TVM_REQUIRE(((uintptr_t)out & 15) == 0, "tc_weights must be 16-byte aligned");
  bgtc::k_pack_bg_tc<<<dim3(16, 5), 256, 0, (cudaStream_t)stream>>>(b, (uint8_t*)out);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
