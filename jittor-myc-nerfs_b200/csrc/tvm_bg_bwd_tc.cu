// Backward of the NeRF++ background network on the 5th-generation tensor cores (tcgen05 / TMEM, sm_100a only).
// Replaces k_bg_bwd (fp32 FMA, tvm_bg_bwd.cu) when the step runs in a tensor-core mode: gradients of MLPNet's parameters
// (models/nerfplusplus.py:66-140) through the 512-sample background compositing (:283-317); Jittor autograd in the reference
// (train.py:228,260).
//
// One persistent CTA per SM, rays handed out by a global counter, tiles of 128 samples in the forward kernel's order (k_bg_tc).
// 256 threads: threads t and t + 128 both own sample row t (= TMEM lane t; warps w and w + 4 share lane quadrant w) and split the
// columns of every epilogue between them; the row's scalar chain (geometry, compositing, its backward) is computed by both.  Per tile every dense product runs as tcgen05.mma with bf16 operands and
// fp32 accumulation; ONE shared-memory image per tensor serves the forward product, the data gradient and the weight
// gradient, read K-major or MN-major ("transposed", tvm_tc_selftest.cu):
//
//   forward recompute (bit-identical to k_bg_tc)   X0 -> Y0 -> Y1 -> Y2 -> (HID | sigma) -> rgb; front-to-back compositing
//   composite backward      dL/d c_j = g' w_j,  dL/d alpha_j = T_j g'.c_j - (g'.C_bg - prefix_j) / (1 - alpha_j + 1e-6)
//                           (C_bg kept by the forward pass, so one forward sweep suffices -- as k_bg_bwd)
//   DH  = (W_rgb^T dlogit) * [hid > 0]                     CUDA cores (rank 3), over HID
//   dWf   (+)= Y2^T DH          A = Y2 (MN), B = DH (MN)                                   persistent in TMEM
//   dY2   = DH Wf               A = DH (K),  B = forward W3 image (MN); + d sigma_pre * w_sigma in the epilogue
//   D2  = dY2 * [y2 > 0]        over Y2
//   dW2h  (+)= Y1^T D2          A = Y1 (MN), B = D2 (MN)                                   persistent in TMEM
//   dW2p^T(+)= D2^T X0          A = D2 (MN), B = X0 (MN): lane = output unit, column = position column (20 = bias b2)   persistent
//   dY1   = D2 W2h              A = D2 (K),  B = forward W2 image, hidden part (MN)
//   D1  = dY1 * [y1 > 0]        over Y1
//   Y0 recomputed (one K = 32 product) into the dead D2 image: keeping it alive would not fit shared memory
//   dW1   (+)= Y0^T D1 ,  dY0 = D1 W1 ,  D0 = dY0 * [y0 > 0] ,  dW0^T (+)= D0^T X0 (column 20 = bias b0)                 persistent
// The persistent accumulators (dW1 128 + dW2h 128 + dWf 64 + dW2p^T 32 + dW0^T 32 TMEM columns behind 128 working columns =
// all 512) are flushed once per CTA.  Column sums and the rank-3 / rank-1 heads (b1, the per-ray bias of the hidden colour
// layer, w_rgb, b_rgb, w_sigma, b_sigma) are accumulated by the row threads in shared memory with conflict-free atomics.
#include "tvm_bg.cuh"
#include "tvm_bwd.cuh"
#include "tvm_tc.cuh"

namespace tvm {
namespace bgbwd {

using namespace tc;
using namespace bgimg;

constexpr int kThreads = 256;                         // two warps per TMEM lane quadrant: each takes half of the columns of every epilogue
constexpr uint32_t kLbo = kRows * 16;                  // 2048: next 8 columns of a 128-row image
// TMEM columns
constexpr uint32_t cWork = 0, cDW1 = 128, cDW2 = 256, cDWf = 384, cDW2p = 448, cDW0 = 480;
constexpr uint32_t MN = kIdescAMajorMN | kIdescBMajorMN;

__device__ __forceinline__ void row_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// Column sums over the 32 rows a warp owns, for 32 columns at once: lane l returns sum_rows d[l].  Butterfly transpose-reduce:
// at distance o a lane keeps the half of its values that matches bit o of its lane index and adds the partner's copy of that
// half -- 31 shuffles instead of 32 x 5.  `d` is destroyed.
__device__ __forceinline__ float warp_colsum32(float (&d)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? d[i] : d[i + o];
      const float keep = up ? d[i + o] : d[i];
      d[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return d[0];
}

// shared-memory carve-up (bytes)
constexpr uint32_t kOffXA = (kImageBytes + 127) & ~127u;         // [128 x 160] bf16: X0 (cols 0..31) | Y0 -> Y1 -> D1 (cols 32..159)
constexpr uint32_t kOffY2 = kOffXA + kRows * kAK * 2;            // [128 x 128]: Y2 -> D2 -> Y0 (recomputed) -> D0
constexpr uint32_t kOffHID = kOffY2 + kRows * kFeatureC * 2;     // [128 x 64]:  HID -> DH
constexpr uint32_t kOffF = kOffHID + kRows * kBgHid * 2;         // fp32 scratch, see below
constexpr int kFVB = 0, kFWRGB = 64, kFWSIG = 256, kFDB1 = 384, kFDVB = 512, kFDWRGB = 576, kFDWSIG = 768, kFSMALL = 896,
              kFWP = 904, kFWQ = 908, kFE = 912, kFEnd = 928;
constexpr uint32_t kOffBar = kOffF + kFEnd * 4;
constexpr uint32_t kSmemBytes = kOffBar + 32;

struct Params {
  FwdParams f;
  const float* d_rgb_map;
  TvmBgGrads g;
};

__global__ void __launch_bounds__(kThreads, 1) k_bg_bwd_tc(const Params Bp) {
  extern __shared__ __align__(128) uint8_t smem[];
  const FwdParams& P = Bp.f;
  const TvmBgNet& bg = P.bg;
  uint8_t* sW = smem;
  uint8_t* sXA = smem + kOffXA;
  uint8_t* sY2 = smem + kOffY2;
  uint8_t* sHID = smem + kOffHID;
  float* F = reinterpret_cast<float*>(smem + kOffF);
  float *VB = F + kFVB, *WRGB = F + kFWRGB, *WSIG = F + kFWSIG, *DB1 = F + kFDB1, *DVB = F + kFDVB, *DWRGB = F + kFDWRGB,
        *DWSIG = F + kFDWSIG, *SMALL = F + kFSMALL, *WP = F + kFWP, *WQ = F + kFWQ, *E = F + kFE;
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  uint32_t* sRay = tmem_slot + 1;

  const int tid = threadIdx.x & 127, half = threadIdx.x >> 7, warp = tid >> 5, lane = tid & 31;    // tid: sample row; half: column half
  const bool lead = half == 0;                                                                     // the half that does the unique writes
  const float R = P.m.radii;

  {
    const uint4* src = reinterpret_cast<const uint4*>(bg.tc_weights);
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (uint32_t i = threadIdx.x; i < kImageBytes / 16; i += kThreads) dst[i] = __ldg(src + i);
  }
  for (int i = threadIdx.x; i < 3 * kBgHid; i += kThreads) WRGB[i] = bg.w_rgb[i];
  if (lead) WSIG[tid] = bg.w_sigma[tid];
  for (int i = threadIdx.x; i < kFEnd - kFDB1; i += kThreads) F[kFDB1 + i] = 0.0f;      // every accumulator (and WP / WQ / E)
  if (threadIdx.x == 0) {
    mbar_init(mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) tmem_alloc(tmem_slot, 512);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t aXA = smem_u32(sXA), aH1 = aXA + (kPosK / 8) * kLbo, aY2 = smem_u32(sY2), aHID = smem_u32(sHID);
  const uint32_t aW0 = smem_u32(sW + kOffW0), aW1 = smem_u32(sW + kOffW1), aW2 = smem_u32(sW + kOffW2),
                 aW3 = smem_u32(sW + kOffW3), aW4 = smem_u32(sW + kOffW4);
  const float b_sigma = reinterpret_cast<const float*>(sW + kOffF32)[0];
  const float b_r = reinterpret_cast<const float*>(sW + kOffF32)[1], b_g = reinterpret_cast<const float*>(sW + kOffF32)[2],
              b_b = reinterpret_cast<const float*>(sW + kOffF32)[3];
  uint8_t* xrow = sXA + tid * 16;                      // this row of the position block
  uint8_t* h1row = xrow + (kPosK / 8) * kLbo;          // ... of the hidden slot (Y0 / Y1 / D1)
  uint8_t* y2row = sY2 + tid * 16;
  uint8_t* hidrow = sHID + tid * 16;
  constexpr uint32_t ID128 = instr_desc(128, 128), ID80 = instr_desc(128, kN3), ID16 = instr_desc(128, kN4),
                     ID64 = instr_desc(128, kBgHid), ID32 = instr_desc(128, kPosK);
  const uint32_t n_active = P.ws.n_entries[1];
  const bool issuer = threadIdx.x == 0;
  uint32_t phase = 0;
  bool first = true;                                   // no persistent accumulator has been written yet

  auto mma_wait = [&]() {
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    fence_after();
  };
  auto publish = [&]() {      // generic-proxy writes of the images -> visible to the tensor core, then CTA barrier
    fence_async_smem();
    fence_before();
    row_sync();
  };
  // K-major product D (+)= A[:, 16 s ..] . B_image^T over `steps` K-steps (forward products and data gradients' A side)
  auto kmajor = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t lbo_b, int steps, uint32_t idesc, bool acc_first) {
    for (int s = 0; s < steps; ++s)
      umma_bf16(d, smem_desc(a + s * 2 * kLbo, kLbo, 128), smem_desc(b + s * 2 * lbo_b, lbo_b, 128), idesc, (acc_first || s > 0) ? 1u : 0u);
  };
  // accumulator columns [0, 128) of this row -> ReLU -> bf16 -> 16 chunks of a 128-row image
  auto epi_relu_store = [&](uint8_t* dst_row) {
#pragma unroll
    for (int cb = 2 * half; cb < 2 * half + 2; ++cb) {
      float y[32];
      tmem_ld32(lane_addr + cWork + cb * 32, y);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(dst_row + (cb * 4 + g) * kLbo) =
            make_uint4(pack_relu_bf16(y[g * 8 + 0], y[g * 8 + 1]), pack_relu_bf16(y[g * 8 + 2], y[g * 8 + 3]),
                       pack_relu_bf16(y[g * 8 + 4], y[g * 8 + 5]), pack_relu_bf16(y[g * 8 + 6], y[g * 8 + 7]));
    }
  };
  // accumulator columns [0, 128) (+ extra * wsig) masked by the sign of the stored activation, in place over that image;
  // optionally adds the column sums of the masked values to a shared accumulator
  auto epi_delta = [&](uint8_t* row_img, float extra, float* colsum) {
#pragma unroll
    for (int cb = 2 * half; cb < 2 * half + 2; ++cb) {
      float d[32];
      tmem_ld32(lane_addr + cWork + cb * 32, d);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4* p = reinterpret_cast<uint4*>(row_img + (cb * 4 + g) * kLbo);
        const uint4 y = *p;
        const uint32_t yy[4] = {y.x, y.y, y.z, y.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = cb * 32 + g * 8 + 2 * i;
          const float lo = bf_lo(yy[i]) > 0.0f ? fmaf(extra, WSIG[c], d[g * 8 + 2 * i]) : 0.0f;
          const float hi = bf_hi(yy[i]) > 0.0f ? fmaf(extra, WSIG[c + 1], d[g * 8 + 2 * i + 1]) : 0.0f;
          d[g * 8 + 2 * i] = lo;
          d[g * 8 + 2 * i + 1] = hi;
          o[i] = pack_bf16(lo, hi);
        }
        *p = make_uint4(o[0], o[1], o[2], o[3]);
      }
      if (colsum) atomicAdd(colsum + cb * 32 + lane, warp_colsum32(d, lane));      // one conflict-free atomic per warp and 32 columns
    }
  };

  while (true) {
    row_sync();
    if (threadIdx.x == 0) *sRay = atomicAdd(P.ws.n_entries + 3, 1u);
    row_sync();
    const uint32_t idx = *sRay;
    if (idx >= n_active) break;
    const uint32_t ray = P.ws.bg_list[idx];
    const float* ray6 = P.rays + 6 * (size_t)ray;
    const float* rnd = P.bg_rand + (size_t)ray * kBgSamples;
    const float lam = P.ws.bg_lambda[ray];
    const float g0 = lam * Bp.d_rgb_map[(size_t)ray * 3 + 0], g1 = lam * Bp.d_rgb_map[(size_t)ray * 3 + 1],
                g2 = lam * Bp.d_rgb_map[(size_t)ray * 3 + 2];
    if (g0 == 0.0f && g1 == 0.0f && g2 == 0.0f) continue;     // uniform over the CTA
    const float gtot = g0 * P.ws.bg_rgb[(size_t)ray * 3 + 0] + g1 * P.ws.bg_rgb[(size_t)ray * 3 + 1] +
                       g2 * P.ws.bg_rgb[(size_t)ray * 3 + 2];
    BgRay g;
    bg_ray_setup(ray6, R, g);
    if (lead && tid < kBgHid) {
      // view-direction embedding through its slice of rgb_layers.0, plus the folded bias (fp32, once per ray; as k_bg_tc)
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
      float a = bg.bf[tid];
#pragma unroll
      for (int j = 0; j < kDirDim; ++j) a = fmaf(e[j], __ldg(bg.wv_t + j * kBgHid + tid), a);
      VB[tid] = a;
      DVB[tid] = 0.0f;
      if (tid == 0) {
#pragma unroll
        for (int j = 0; j < kDirDim; ++j) E[j] = e[j];
      }
    }
    float T = 1.0f, carry = 0.0f;

#pragma unroll 1
    for (int tile = 0; tile < kBgSamples / kRows; ++tile) {
      const uint32_t acc_flag = first ? 0u : 1u;
      // ---- geometry + embedding of this row's sample (flipped order j, original index i = 511 - j; as k_bg_tc) -----------
      float dz;
      {
        const int j = tile * kRows + tid, i = kBgSamples - 1 - j;
        const float z = bg_depth(i, R, rnd);
        dz = (i > 0) ? z - bg_depth(i - 1, R, rnd) : 1e10f;             // bg_dists, HUGE_NUMBER last (:299-300)
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
        if (lead) {
        *reinterpret_cast<uint4*>(xrow + 0 * kLbo) = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(s1[0], s1[1]), pack_bf16(s1[2], s1[3]));
        *reinterpret_cast<uint4*>(xrow + 1 * kLbo) = make_uint4(pack_bf16(cc1[0], cc1[1]), pack_bf16(cc1[2], cc1[3]), pack_bf16(s2[0], s2[1]), pack_bf16(s2[2], s2[3]));
        *reinterpret_cast<uint4*>(xrow + 2 * kLbo) = make_uint4(pack_bf16(cc2[0], cc2[1]), pack_bf16(cc2[2], cc2[3]), pack_bf16(1.0f, 0.0f), 0u);
        *reinterpret_cast<uint4*>(xrow + 3 * kLbo) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      publish();
      // ================================ forward recompute =======================================================
      if (issuer) {                                               // L0: X0 -> Y0
        fence_after();
        kmajor(tmem + cWork, aXA, aW0, kFeatureC * 16, kPosK / 16, ID128, false);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_relu_store(h1row);
      publish();
      if (issuer) {                                               // L1: Y0 (+ bias through position columns 16..31) -> Y1
        fence_after();
        kmajor(tmem + cWork, aH1, aW1, kFeatureC * 16, kFeatureC / 16, ID128, false);
        kmajor(tmem + cWork, aXA + 2 * kLbo, aW1 + (kFeatureC / 8) * kFeatureC * 16, kFeatureC * 16, 1, ID128, true);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_relu_store(h1row);
      publish();
      if (issuer) {                                               // L2: [X0 | Y1] -> Y2
        fence_after();
        kmajor(tmem + cWork, aXA, aW2, kFeatureC * 16, kAK / 16, ID128, false);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_relu_store(y2row);
      publish();
      if (issuer) {                                               // L3: Y2 -> (HID | sigma)
        fence_after();
        kmajor(tmem + cWork, aY2, aW3, kN3 * 16, kFeatureC / 16, ID80, false);
        umma_commit(mma_bar);
      }
      mma_wait();
      float sig_pre;
      {
        {
          const int cb = half;
          float y[32];
          tmem_ld32(lane_addr + cWork + cb * 32, y);
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const float4 b = *reinterpret_cast<const float4*>(VB + cb * 32 + q);
            y[q] += b.x; y[q + 1] += b.y; y[q + 2] += b.z; y[q + 3] += b.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(hidrow + (cb * 4 + q) * kLbo) =
                make_uint4(pack_relu_bf16(y[q * 8 + 0], y[q * 8 + 1]), pack_relu_bf16(y[q * 8 + 2], y[q * 8 + 3]),
                           pack_relu_bf16(y[q * 8 + 4], y[q * 8 + 5]), pack_relu_bf16(y[q * 8 + 6], y[q * 8 + 7]));
        }
        float t16[16];
        tmem_ld16(lane_addr + cWork + 64, t16);
        sig_pre = t16[0] + b_sigma;                                          // sigma = |w . base + b| (:128-129)
      }
      publish();
      if (issuer) {                                               // L4: HID -> logits
        fence_after();
        kmajor(tmem + cWork, aHID, aW4, kN4 * 16, kBgHid / 16, ID16, false);
        umma_commit(mma_bar);
      }
      mma_wait();
      float rgb[3];
      {
        float t16[16];
        tmem_ld16(lane_addr + cWork, t16);
        rgb[0] = 1.0f / (1.0f + __expf(-(t16[0] + b_r)));
        rgb[1] = 1.0f / (1.0f + __expf(-(t16[1] + b_g)));
        rgb[2] = 1.0f / (1.0f + __expf(-(t16[2] + b_b)));
      }
      // ================================ compositing and its backward ===========================================
      const float sigma = fabsf(sig_pre);
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
      if (lead && lane == 31) WP[warp] = pref;
      fence_before();
      row_sync();
      const float p0 = WP[0], p1 = WP[1], p2 = WP[2], p3 = WP[3];
      const float before = warp == 0 ? 1.0f : warp == 1 ? p0 : warp == 2 ? p0 * p1 : p0 * p1 * p2;
      const float Tj = T * before * excl;
      const float w = al * Tj;
      const float gc = g0 * rgb[0] + g1 * rgb[1] + g2 * rgb[2];             // g' . c_j
      float ps = w * gc;                                                    // inclusive prefix of w_j g'.c_j in sample order
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, ps, o);
        if (lane >= o) ps += t;
      }
      if (lead && lane == 31) WQ[warp] = ps;
      row_sync();
      const float q0 = WQ[0], q1 = WQ[1], q2 = WQ[2], q3 = WQ[3];
      const float incl = carry + ps + (warp == 0 ? 0.0f : warp == 1 ? q0 : warp == 2 ? q0 + q1 : q0 + q1 + q2);
      const float dalpha = Tj * gc - (gtot - incl) / vv;
      const float dsig = dalpha * dz * (1.0f - al);                         // alpha = 1 - exp(-sigma dist)
      const float dsp = dsig * (sig_pre > 0.0f ? 1.0f : (sig_pre < 0.0f ? -1.0f : 0.0f));     // through |.|
      const float dl0 = g0 * w * rgb[0] * (1.0f - rgb[0]), dl1 = g1 * w * rgb[1] * (1.0f - rgb[1]),
                  dl2 = g2 * w * rgb[2] * (1.0f - rgb[2]);
      carry += (q0 + q1) + (q2 + q3);
      T = T * (p0 * p1 * p2 * p3);
      {
        // b_rgb, b_sigma
        const float s0 = warp_sum(dl0), s1 = warp_sum(dl1), s2 = warp_sum(dl2), s3 = warp_sum(dsp);
        if (lead && lane == 0) {
          atomicAdd(SMALL + 0, s0);
          atomicAdd(SMALL + 1, s1);
          atomicAdd(SMALL + 2, s2);
          atomicAdd(SMALL + 3, s3);
        }
      }
      // ---- rank-3 colour layer: d w_rgb += dlogit (x) hid;  DH = (W_rgb^T dlogit) * [hid > 0] over HID;  d VB += DH ---------
      // ---- rank-1 sigma head:   d w_sigma += d sigma_pre * y2 -----------------------------------------------------------------
      {
        const int cb = half;
        float h[32], dh[32];
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          uint4* p = reinterpret_cast<uint4*>(hidrow + (cb * 4 + kc) * kLbo);
          const uint4 hv = *p;
          const uint32_t hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { h[kc * 8 + 2 * i] = bf_lo(hh[i]); h[kc * 8 + 2 * i + 1] = bf_hi(hh[i]); }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = cb * 32 + kc * 8 + i;
            dh[kc * 8 + i] = h[kc * 8 + i] > 0.0f ? fmaf(dl0, WRGB[c], fmaf(dl1, WRGB[kBgHid + c], dl2 * WRGB[2 * kBgHid + c])) : 0.0f;
          }
          *p = make_uint4(pack_bf16(dh[kc * 8 + 0], dh[kc * 8 + 1]), pack_bf16(dh[kc * 8 + 2], dh[kc * 8 + 3]),
                          pack_bf16(dh[kc * 8 + 4], dh[kc * 8 + 5]), pack_bf16(dh[kc * 8 + 6], dh[kc * 8 + 7]));
        }
        float t[32];
        atomicAdd(DVB + cb * 32 + lane, warp_colsum32(dh, lane));
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = dl0 * h[i];
        atomicAdd(DWRGB + cb * 32 + lane, warp_colsum32(t, lane));
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = dl1 * h[i];
        atomicAdd(DWRGB + kBgHid + cb * 32 + lane, warp_colsum32(t, lane));
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = dl2 * h[i];
        atomicAdd(DWRGB + 2 * kBgHid + cb * 32 + lane, warp_colsum32(t, lane));
      }
#pragma unroll 1
      for (int cb = 2 * half; cb < 2 * half + 2; ++cb) {
        float t[32];
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          const uint4 yv = *reinterpret_cast<const uint4*>(y2row + (cb * 4 + kc) * kLbo);
          const uint32_t yy[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { t[kc * 8 + 2 * i] = dsp * bf_lo(yy[i]); t[kc * 8 + 2 * i + 1] = dsp * bf_hi(yy[i]); }
        }
        atomicAdd(DWSIG + cb * 32 + lane, warp_colsum32(t, lane));
      }
      publish();
      // ================================ dense backward ================================================================
      if (issuer) {
        fence_after();
        // dWf (+)= Y2^T DH   [128 y2 features x 64 hidden units], reduction over the 128 rows
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cDWf, smem_desc(aY2 + s * 256, 128, kLbo), smem_desc(aHID + s * 256, 128, kLbo), ID64 | MN, acc_flag | (s > 0));
        // dY2 = DH Wf   (forward W3 image read transposed: reduction over its 64 colour-layer rows, 80 rows per K-chunk)
        for (int s = 0; s < kBgHid / 16; ++s)
          umma_bf16(tmem + cWork, smem_desc(aHID + s * 2 * kLbo, kLbo, 128), smem_desc(aW3 + s * 256, 128, kN3 * 16),
                    ID128 | kIdescBMajorMN, s > 0);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_delta(y2row, dsp, nullptr);                             // D2 = (dY2 + d sigma_pre w_sigma) * [y2 > 0], over Y2
      publish();
      if (issuer) {
        fence_after();
        for (int s = 0; s < 8; ++s)                               // dW2h (+)= Y1^T D2
          umma_bf16(tmem + cDW2, smem_desc(aH1 + s * 256, 128, kLbo), smem_desc(aY2 + s * 256, 128, kLbo), ID128 | MN, acc_flag | (s > 0));
        for (int s = 0; s < 8; ++s)                               // dW2p^T (+)= D2^T X0   [128 output units x 32 position columns]
          umma_bf16(tmem + cDW2p, smem_desc(aY2 + s * 256, 128, kLbo), smem_desc(aXA + s * 256, 128, kLbo), ID32 | MN, acc_flag | (s > 0));
        for (int s = 0; s < 8; ++s)                               // dY1 = D2 W2h (hidden part of the W2 image: K-chunks 4..19)
          umma_bf16(tmem + cWork, smem_desc(aY2 + s * 2 * kLbo, kLbo, 128), smem_desc(aW2 + (kPosK / 8) * kLbo + s * 256, 128, kLbo),
                    ID128 | kIdescBMajorMN, s > 0);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_delta(h1row, 0.0f, DB1);                                // D1 = dY1 * [y1 > 0], over Y1; b1 += column sums
      publish();
      if (issuer) {                                               // Y0 again (it was overwritten by Y1): X0 -> work
        fence_after();
        kmajor(tmem + cWork, aXA, aW0, kFeatureC * 16, kPosK / 16, ID128, false);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_relu_store(y2row);                                      // into the dead D2 image
      publish();
      if (issuer) {
        fence_after();
        for (int s = 0; s < 8; ++s)                               // dW1 (+)= Y0^T D1
          umma_bf16(tmem + cDW1, smem_desc(aY2 + s * 256, 128, kLbo), smem_desc(aH1 + s * 256, 128, kLbo), ID128 | MN, acc_flag | (s > 0));
        for (int s = 0; s < 8; ++s)                               // dY0 = D1 W1 (hidden rows of the W1 image)
          umma_bf16(tmem + cWork, smem_desc(aH1 + s * 2 * kLbo, kLbo, 128), smem_desc(aW1 + s * 256, 128, kLbo),
                    ID128 | kIdescBMajorMN, s > 0);
        umma_commit(mma_bar);
      }
      mma_wait();
      epi_delta(y2row, 0.0f, nullptr);                            // D0 = dY0 * [y0 > 0], over Y0
      publish();
      if (issuer) {                                               // dW0^T (+)= D0^T X0
        fence_after();
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cDW0, smem_desc(aY2 + s * 256, 128, kLbo), smem_desc(aXA + s * 256, 128, kLbo), ID32 | MN, acc_flag | (s > 0));
        umma_commit(mma_bar);
      }
      mma_wait();                                                 // the images are free for the next tile
      first = false;
      if (T < 1e-6f) break;            // uniform: the forward stopped here as well (remaining weight mass < 1e-6)
    }
    // VB = bf + wv^T e: d bf, d wv_t
    row_sync();
    if (lead && tid < kBgHid) {
      const float d = DVB[tid];
      atomicAdd(Bp.g.bf + tid, d);
#pragma unroll
      for (int j = 0; j < kDirDim; ++j) atomicAdd(Bp.g.wv_t + j * kBgHid + tid, E[j] * d);
    }
  }

  // ================================ flush: persistent accumulators and shared sums ======================================
  row_sync();
  if (!first && lead) {
    fence_after();
    float v[32];
    // dW1: lane = Y0 feature j, column = output unit n -> w1_t[j][n];  dW2h: lane = Y1 feature -> w2_t[20 + j][n]
#pragma unroll 1
    for (int cb = 0; cb < 4; ++cb) {
      tmem_ld32(lane_addr + cDW1 + cb * 32, v);
      float* o1 = Bp.g.w1_t + (size_t)tid * kFeatureC + cb * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) red_add_v4(o1 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
      tmem_ld32(lane_addr + cDW2 + cb * 32, v);
      float* o2 = Bp.g.w2_t + (size_t)(kPosDim + tid) * kFeatureC + cb * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) red_add_v4(o2 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    // dWf: lane = Y2 feature j, column = hidden colour unit h -> wf_t[j][h]
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      tmem_ld32(lane_addr + cDWf + cb * 32, v);
      float* o = Bp.g.wf_t + (size_t)tid * kBgHid + cb * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) red_add_v4(o + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    // dW2p^T / dW0^T: lane = output unit n, column = position column k (k < 20: weight row k; k = 20: the bias)
    tmem_ld32(lane_addr + cDW2p, v);
#pragma unroll
    for (int k = 0; k < kPosDim; ++k) atomicAdd(Bp.g.w2_t + (size_t)k * kFeatureC + tid, v[k]);
    atomicAdd(Bp.g.b2 + tid, v[kOneCol]);
    tmem_ld32(lane_addr + cDW0, v);
#pragma unroll
    for (int k = 0; k < kPosDim; ++k) atomicAdd(Bp.g.w0_t + (size_t)k * kFeatureC + tid, v[k]);
    atomicAdd(Bp.g.b0 + tid, v[kOneCol]);
    atomicAdd(Bp.g.b1 + tid, DB1[tid]);
    atomicAdd(Bp.g.w_sigma + tid, DWSIG[tid]);
    for (int i = tid; i < 3 * kBgHid; i += 128) atomicAdd(Bp.g.w_rgb + i, DWRGB[i]);
    if (tid < 3) atomicAdd(Bp.g.b_rgb + tid, SMALL[tid]);
    if (tid == 3) atomicAdd(Bp.g.b_sigma, SMALL[3]);
  }
  fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

}  // namespace bgbwd

int launch_bg_bwd_tc(const BwdParams& Bw, const TvmBgGrads& gr, int num_sms, cudaStream_t stream) {
  using namespace bgbwd;
  TVM_REQUIRE(Bw.f.bg.tc_weights != nullptr, "TvmBgNet.tc_weights is NULL: call tvm_pack_bg_tc first");
  static_assert(kSmemBytes <= 227 * 1024, "k_bg_bwd_tc shared memory");
  Params B;
  B.f = Bw.f;
  B.d_rgb_map = Bw.d_rgb_map;
  B.g = gr;
  TVM_CHECK_CUDA(cudaFuncSetAttribute(k_bg_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  k_bg_bwd_tc<<<num_sms, kThreads, kSmemBytes, stream>>>(B);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvm
