// Appearance-head backward on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
// Replaces k_app_bwd (fp32 FMA) when the step runs with TVM_MLP_BF16; row a12 of SURVEY.md §8a for the
// basis_mat + MLPRender_Fea part (tensorf-myc/models/tensoRF.py:228-244, models/tensorBase.py:62-86; gradients
// come from Jittor autograd in the reference, train.py:228,260).
//
// One persistent CTA per SM, 12 warps, tiles of 128 weighted samples.  Every dense product of the step runs as
// tcgen05.mma with bf16 operands and fp32 accumulation in TMEM; ONE shared-memory image per tensor serves the
// forward product, the data gradient and the weight gradient, read K-major or MN-major ("transposed", see
// tvm_tc_selftest.cu):
//
//   forward recompute   H (gather) -> feat = H basis^T -> X = [feat, dir, PE] -> Y1 = relu(X W1^T + b1) -> Y2 = relu(Y1 W2^T + b2)
//                       rgb = sigmoid(W3 y2 + b3);  dO = w * g_ray * rgb (1 - rgb)
//   dW3  (+)= Y2^T dO            A = Y2 (MN), B = dO (MN, parked in the padding columns of X)        persistent in TMEM
//   G2   = (dO W3) * [y2 > 0]    CUDA cores (rank 3)
//   dY1  = G2 W2                 A = G2 (K),  B = forward W2 image (MN)
//   dW2  (+)= G2^T Y1            A = G2 (MN), B = Y1 (MN)                                            persistent in TMEM
//   G1   = dY1 * [y1 > 0]        written over Y1
//   dX   = G1 W1                 A = G1 (K),  B = forward W1 image (MN)
//   dW1  (+)= G1^T X             A = G1 (MN), B = X (MN)                                             persistent in TMEM
//   dF   = positional-encoding backward of dX (CUDA cores), written over X[:, 0:32)
//   dH   = dF basis              A = dF (K),  B = forward basis image (MN)
//   dBasis^T = H^T dF            A = H (MN, two overlapping 128-row windows of the 144 channels), B = dF (MN); flushed per tile
//   (REFTensoRF: dF carries the 48 stacked head gradients -- features, normal, diffuse, specular -- after the reflection /
//    dot-product / normalisation / penalty backward on CUDA cores)
//   scatter dH into the appearance planes / lines with red.global.add.v4.f32 (all 12 warps, 4 lanes per sample)
//
// The persistent accumulators (dW2 128 + dW1 160 + dW3 16 TMEM columns, after 160 working columns) are flushed once per CTA.
#include "tvm_bwd.cuh"
#include "tvm_tc.cuh"

namespace tvm {
namespace bwdtc {

using namespace tc;

constexpr int kRowWarps = 4, kHelperWarps = 8;
constexpr int kThreads = (kRowWarps + kHelperWarps) * 32;
constexpr int CA = 48, APP_DIM = 27, K0 = 3 * CA, K1 = 160;
constexpr int NF = 2 * APP_DIM, NV = 6;                 // PE widths (fea_pe = view_pe = 2)
constexpr int kDoChunk = 19, kDoCol = kDoChunk * 8;     // dO parked in X columns 152..154 (padding of the 150/151 real columns)

constexpr uint32_t kLbo = kRows * 16;                  // 2048: next 8 columns of a 128-row image
// TMEM columns: [0,160) working accumulators (forward layers, dY1, dX, dH, dBasis^T windows), then the persistent dW2, dW1, dW3
constexpr uint32_t cWork = 0, cBasis = 0, cDW2 = 160, cDW1 = 288, cDW3 = 448;

__device__ __forceinline__ void row_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// REF = REFTensoRF (models/REFTensoRF.py): 48 stacked head outputs (basis | normal | diffuse | specular | rho), MLP input
// [-d.n, feat, reflection, PE(feat), PE(reflection)], rgb = tint * rgb_s + rgb_d, normal penalty.
template <bool REF>
__global__ void __launch_bounds__(kThreads, 1) k_app_bwd_tc(const BwdParams Bp) {
  constexpr int NH = REF ? TVM_REF_HEAD_LD : 32;
  constexpr int C0 = REF ? 1 : 0;                                      // REF: column 0 of X is -d.n
  constexpr int IN_C = 150 + C0;
  constexpr int PF = C0 + APP_DIM + 3, PV = PF + 2 * NF;               // column map of X (tensorBase.py:76-83, REFTensoRF.py:19-25)
  constexpr uint32_t cDBt1 = 0, cDBt2 = NH;
  extern __shared__ __align__(128) uint8_t smem[];
  const FwdParams& P = Bp.f;
  const TvmModel& m = P.m;
  const Image img(CA, IN_C, NH);
  // REF: the 13.8 KB head image leaves no room for the fp32 tail (biases, W3) in shared memory: it is read through L1
  const uint32_t w_bytes = REF ? img.off_f32 : img.bytes;
  uint8_t* sW = smem;
  uint8_t* sH = smem + ((w_bytes + 127) & ~127u);        // [128 x 144] bf16: H, later dH
  uint8_t* sX = sH + kRows * K0 * 2;                     // [128 x 160] bf16: X (+ dO), later dF in columns 0..31
  uint8_t* sY1 = sX + kRows * K1 * 2;                    // [128 x 128] bf16: Y1, later G1
  uint8_t* sG2 = sY1 + kRows * 128 * 2;                  // [128 x 128] bf16: Y2, later G2
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sG2 + kRows * 128 * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const float* sB1 = REF ? reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(m.tc_weights_bwd ? m.tc_weights_bwd : m.tc_weights) + img.off_f32)
                         : reinterpret_cast<const float*>(sW + img.off_f32);
  const float* sB2 = sB1 + 128;
  const float* sW3 = sB2 + 128;
  const float* sB3 = sW3 + 3 * 128;
  const float* sHB = sB3 + 4;                            // REF: biases of the stacked heads [48]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const uint4* src = reinterpret_cast<const uint4*>(m.tc_weights_bwd ? m.tc_weights_bwd : m.tc_weights);   // bf16 image
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (uint32_t i = tid; i < w_bytes / 16; i += kThreads) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t aH = smem_u32(sH), aX = smem_u32(sX), aY1 = smem_u32(sY1), aG2 = smem_u32(sG2);
  const uint32_t aB0 = smem_u32(sW + img.off_b0), aB1 = smem_u32(sW + img.off_b1), aB2 = smem_u32(sW + img.off_b2);
  constexpr uint32_t MN = kIdescAMajorMN | kIdescBMajorMN;

  // the row warps carry the big epilogues; the helper warps only gather / scatter
  if (warp < kRowWarps) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 128;");

  const uint32_t n_ent_all = *P.ws.n_entries;
  const uint32_t tile_rows = balanced_tile_rows(n_ent_all, gridDim.x);       // 128, or fewer for launches of a few rounds
  const uint32_t n_tiles = (n_ent_all + tile_rows - 1) / tile_rows;
  const float4* gray = reinterpret_cast<const float4*>(P.ws.bwd_scratch);
  uint32_t phase = 0;
  bool first = true;
  float db3[3] = {0.0f, 0.0f, 0.0f};     // row threads: sum of dO over their rows
  float dhb[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // REF row threads: sum of the head-output gradients
  const float dpen = (REF && Bp.d_penalty) ? *Bp.d_penalty : 0.0f;
  float dbias = 0.0f;                    // helper threads: column sum of G2 (warps 4-7) / G1 (warps 8-11)

  auto mma_wait = [&]() {
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    fence_after();
  };
  auto publish = [&]() {
    fence_async_smem();
    fence_before();
    row_sync();
  };

  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t tile_base = tile * tile_rows;
    const uint32_t n_ent = min(n_ent_all, tile_base + tile_rows);            // rows behind the tile's entries are dead
    // ================================ gather (helper warps) ==============================================
    float dir[3] = {0.0f, 0.0f, 0.0f}, gw[3] = {0.0f, 0.0f, 0.0f}, wgt = 0.0f;
    if (warp >= kRowWarps) {
      const int gwi = warp - kRowWarps;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int row = gwi * 16 + pass * 8 + (lane >> 2), q = lane & 3;
        const uint32_t e = tile_base + row;
        uint8_t* arow = sH + row * 16;
        if (e < n_ent) {
          const float4 uw = __ldg(P.ws.ent_u + e);      // grid coordinates stored by k_march (the forward's workspace)
          const float u[3] = {uw.x, uw.y, uw.z};
          Axis ax[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) ax[i] = axis_taps(u[i], m.grid[i]);
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const VmTaps t = vm_taps(m, ax, kk);
#pragma unroll
            for (int c = q * 4; c < CA; c += 16) {
              float4 pv, lv;
              vm_sample4(m.app_plane[kk], m.app_line[kk], t, CA, c, pv, lv);
              const int k = kk * CA + c;
              *reinterpret_cast<uint2*>(arow + (k >> 3) * kLbo + (k & 7) * 2) =
                  make_uint2(pack_bf16(pv.x * lv.x, pv.y * lv.y), pack_bf16(pv.z * lv.z, pv.w * lv.w));
            }
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk)
#pragma unroll
            for (int c = q * 4; c < CA; c += 16) {
              const int k = kk * CA + c;
              *reinterpret_cast<uint2*>(arow + (k >> 3) * kLbo + (k & 7) * 2) = make_uint2(0u, 0u);
            }
        }
      }
    } else {
      const uint32_t e = tile_base + tid;
      if (e < n_ent) {
        const uint32_t ray = P.ws.ent[e].x;
        dir[0] = P.rays[6 * (size_t)ray + 3];
        dir[1] = P.rays[6 * (size_t)ray + 4];
        dir[2] = P.rays[6 * (size_t)ray + 5];
        const float4 g = gray[ray];
        wgt = P.ws.ent_w[e];
        gw[0] = wgt * g.x; gw[1] = wgt * g.y; gw[2] = wgt * g.z;      // d (w rgb . g) / d rgb
      }
    }
    fence_async_smem();
    fence_before();
    __syncthreads();

    if (warp < kRowWarps) {
      // ================================ row warps: the dense chain ===================================
      const int row = tid;
      uint8_t* xrow = sX + row * 16;
      uint8_t* y1row = sY1 + row * 16;
      uint8_t* g2row = sG2 + row * 16;
      const uint32_t acc_flag = first ? 0u : 1u;
      // ---- GEMM0: feat = H basis^T ------------------------------------------------------------------
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int k = 0; k < K0 / 16; ++k)
          umma_bf16(tmem + cBasis, smem_desc(aH + k * 2 * kLbo, kLbo, 128), smem_desc(aB0 + k * 2 * (NH * 16), NH * 16, 128),
                    instr_desc(128, NH), k > 0);
        umma_commit(mma_bar);
      }
      mma_wait();
      // REF geometry kept for the backward: view d = -dir, raw normal, unit normal, d.n, tint
      float vd[3] = {-dir[0], -dir[1], -dir[2]}, nraw2 = 0.0f, ninv = 0.0f, nh[3] = {0.0f, 0.0f, 0.0f}, dotp = 0.0f, tint = 1.0f;
      {
        // ---- epi0: X = [(-d.n), feat, dir | reflection, sin/cos PE] (tensorBase.py:76-83, 9-15; REFTensoRF.py:216-232) --
        float x[32];
        tmem_ld32(lane_addr + cBasis, x);
        float ndot = 0.0f;
        if (REF) {
          float hx[16];
          tmem_ld16(lane_addr + cBasis + 32, hx);
          auto head = [&](int o) { return (o < 32 ? x[o] : hx[o - 32]) + sHB[o]; };
          const float v0 = head(APP_DIM), v1 = head(APP_DIM + 1), v2 = head(APP_DIM + 2);
          tint = fmaxf(head(APP_DIM + 6), 0.0f);
          nraw2 = v0 * v0 + v1 * v1 + v2 * v2;
          ninv = rsqrtf(fmaxf(nraw2, 1e-30f));
          nh[0] = v0 * ninv; nh[1] = v1 * ninv; nh[2] = v2 * ninv;
          dotp = vd[0] * nh[0] + vd[1] * nh[1] + vd[2] * nh[2];
#pragma unroll
          for (int c = 0; c < 3; ++c) dir[c] = 2.0f * dotp * nh[c] - vd[c];      // reflection replaces the view direction
          ndot = -dotp;
        }
        float s1[APP_DIM + 3], c1[APP_DIM + 3];
#pragma unroll
        for (int o = 0; o < APP_DIM; ++o) __sincosf(x[o], &s1[o], &c1[o]);
#pragma unroll
        for (int o = 0; o < 3; ++o) __sincosf(dir[o], &s1[APP_DIM + o], &c1[APP_DIM + o]);
        auto column = [&](int cc) -> float {
          if (REF && cc == 0) return ndot;
          const int c = cc - C0;
          constexpr int pf = APP_DIM + 3, pv = pf + 2 * NF;
          if (c < APP_DIM) return x[c];
          if (c < pf) return dir[c - APP_DIM];
          if (c < pf + NF) { const int o = (c - pf) >> 1; return ((c - pf) & 1) ? 2.0f * s1[o] * c1[o] : s1[o]; }
          if (c < pv) { const int o = (c - pf - NF) >> 1; return ((c - pf - NF) & 1) ? 1.0f - 2.0f * s1[o] * s1[o] : c1[o]; }
          if (c < pv + NV) { const int o = APP_DIM + ((c - pv) >> 1); return ((c - pv) & 1) ? 2.0f * s1[o] * c1[o] : s1[o]; }
          if (c < pv + 2 * NV) { const int o = APP_DIM + ((c - pv - NV) >> 1); return ((c - pv - NV) & 1) ? 1.0f - 2.0f * s1[o] * s1[o] : c1[o]; }
          return 0.0f;
        };
#pragma unroll
        for (int kc = 0; kc < K1 / 8; ++kc) {
          uint4 v;
          v.x = pack_bf16(column(kc * 8 + 0), column(kc * 8 + 1));
          v.y = pack_bf16(column(kc * 8 + 2), column(kc * 8 + 3));
          v.z = pack_bf16(column(kc * 8 + 4), column(kc * 8 + 5));
          v.w = pack_bf16(column(kc * 8 + 6), column(kc * 8 + 7));
          *reinterpret_cast<uint4*>(xrow + kc * kLbo) = v;
        }
      }
      publish();
      // ---- GEMM1: Y1 = relu(X W1^T + b1) ---------------------------------------------------------------
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int k = 0; k < K1 / 16; ++k)
          umma_bf16(tmem + cWork, smem_desc(aX + k * 2 * kLbo, kLbo, 128), smem_desc(aB1 + k * 2 * kLbo, kLbo, 128),
                    instr_desc(128, 128), k > 0);
        umma_commit(mma_bar);
      }
      mma_wait();
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float y[32];
        tmem_ld32(lane_addr + cWork + cb * 32, y);
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j] + sB1[cb * 32 + j], 0.0f);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(y1row + (cb * 4 + g) * kLbo) =
              make_uint4(pack_bf16(y[g * 8 + 0], y[g * 8 + 1]), pack_bf16(y[g * 8 + 2], y[g * 8 + 3]),
                         pack_bf16(y[g * 8 + 4], y[g * 8 + 5]), pack_bf16(y[g * 8 + 6], y[g * 8 + 7]));
      }
      publish();
      // ---- GEMM2: Y2 = relu(Y1 W2^T + b2); rgb; dO ----------------------------------------------------------
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem + cWork, smem_desc(aY1 + k * 2 * kLbo, kLbo, 128), smem_desc(aB2 + k * 2 * kLbo, kLbo, 128),
                    instr_desc(128, 128), k > 0);
        umma_commit(mma_bar);
      }
      mma_wait();
      uint32_t relu2[4];
      float dO[3], dtint = 0.0f;
      {
        float o0 = sB3[0], o1 = sB3[1], o2 = sB3[2];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          float y[32];
          tmem_ld32(lane_addr + cWork + cb * 32, y);
          uint32_t bits = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float h = fmaxf(y[j] + sB2[cb * 32 + j], 0.0f);
            y[j] = h;
            bits |= (h > 0.0f ? 1u : 0u) << j;
            o0 = fmaf(h, sW3[cb * 32 + j], o0);
            o1 = fmaf(h, sW3[128 + cb * 32 + j], o1);
            o2 = fmaf(h, sW3[256 + cb * 32 + j], o2);
          }
          relu2[cb] = bits;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(g2row + (cb * 4 + g) * kLbo) =
                make_uint4(pack_bf16(y[g * 8 + 0], y[g * 8 + 1]), pack_bf16(y[g * 8 + 2], y[g * 8 + 3]),
                           pack_bf16(y[g * 8 + 4], y[g * 8 + 5]), pack_bf16(y[g * 8 + 6], y[g * 8 + 7]));
        }
        const float r0 = 1.0f / (1.0f + __expf(-o0)), r1 = 1.0f / (1.0f + __expf(-o1)), r2 = 1.0f / (1.0f + __expf(-o2));
        // rgb = tint * rgb_s + rgb_d (REF; tint = 1, rgb_d = 0 otherwise); gw = 0 on padding rows
        dO[0] = gw[0] * tint * r0 * (1.0f - r0);       // d/d logit of w * rgb . g
        dO[1] = gw[1] * tint * r1 * (1.0f - r1);
        dO[2] = gw[2] * tint * r2 * (1.0f - r2);
        if (REF) dtint = gw[0] * r0 + gw[1] * r1 + gw[2] * r2;
        db3[0] += dO[0]; db3[1] += dO[1]; db3[2] += dO[2];
        *reinterpret_cast<uint4*>(xrow + kDoChunk * kLbo) = make_uint4(pack_bf16(dO[0], dO[1]), pack_bf16(dO[2], 0.0f), 0u, 0u);
      }
      publish();
      // ---- dW3 (+)= Y2^T dO  (N = 16: X columns 144..159; only columns 152..154 are meaningful) ----------------
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cDW3, smem_desc(aG2 + s * 256, 128, kLbo), smem_desc(aX + (kDoChunk - 1) * kLbo + s * 256, 128, kLbo),
                    instr_desc(128, 16) | MN, acc_flag | (s > 0));
        umma_commit(mma_bar);
      }
      mma_wait();
      // ---- G2 = (dO W3) * [y2 > 0], over Y2 ------------------------------------------------------------------------
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const uint32_t bits = relu2[cb];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = cb * 32 + g * 8 + i;
            const float t = fmaf(dO[0], sW3[c], fmaf(dO[1], sW3[128 + c], dO[2] * sW3[256 + c]));
            v[i] = ((bits >> (g * 8 + i)) & 1u) ? t : 0.0f;
          }
          *reinterpret_cast<uint4*>(g2row + (cb * 4 + g) * kLbo) =
              make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
      }
      publish();
      // ---- dY1 = G2 W2 ; dW2 (+)= G2^T Y1 -----------------------------------------------------------------------------
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cWork, smem_desc(aG2 + s * 2 * kLbo, kLbo, 128), smem_desc(aB2 + s * 256, 128, kLbo),
                    instr_desc(128, 128) | kIdescBMajorMN, s > 0);
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cDW2, smem_desc(aG2 + s * 256, 128, kLbo), smem_desc(aY1 + s * 256, 128, kLbo),
                    instr_desc(128, 128) | MN, acc_flag | (s > 0));
        umma_commit(mma_bar);
      }
      mma_wait();
      // ---- G1 = dY1 * [y1 > 0], over Y1 ----------------------------------------------------------------------------------
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float d[32];
        tmem_ld32(lane_addr + cWork + cb * 32, d);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4* p = reinterpret_cast<uint4*>(y1row + (cb * 4 + g) * kLbo);
          const uint4 y = *p;
          const uint32_t yy[4] = {y.x, y.y, y.z, y.w};
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float lo = bf_lo(yy[i]) > 0.0f ? d[g * 8 + 2 * i] : 0.0f;
            const float hi = bf_hi(yy[i]) > 0.0f ? d[g * 8 + 2 * i + 1] : 0.0f;
            o[i] = pack_bf16(lo, hi);
          }
          *p = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      publish();
      // ---- dX = G1 W1 ; dW1 (+)= G1^T X -------------------------------------------------------------------------------------
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cWork, smem_desc(aY1 + s * 2 * kLbo, kLbo, 128), smem_desc(aB1 + s * 256, 128, kLbo),
                    instr_desc(128, K1) | kIdescBMajorMN, s > 0);
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cDW1, smem_desc(aY1 + s * 256, 128, kLbo), smem_desc(aX + s * 256, 128, kLbo),
                    instr_desc(128, K1) | MN, acc_flag | (s > 0));
        umma_commit(mma_bar);
      }
      mma_wait();
      {
        // ---- positional-encoding backward: d feat_o = dX[feat_o] + sum_q 2^q (dX[sin_oq] cos_oq - dX[cos_oq] sin_oq) ----------
        // (REF: the same for the three reflection components, then the geometry of REFTensoRF.py:216-238 backwards)
        float df[NH];                                   // gradient of the stacked head outputs (VM: features only)
#pragma unroll
        for (int j = 0; j < NH; ++j) df[j] = 0.0f;
        float dr[3] = {0.0f, 0.0f, 0.0f}, dx0 = 0.0f;   // REF: d reflection, d(-d.n)
        // X values of this row, as stored (bf16): 20 chunks of 8 columns
        auto xval = [&](int c) -> float {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(xrow + (c >> 3) * kLbo + (c & 7) / 2 * 4);
          return (c & 1) ? bf_hi(w) : bf_lo(w);
        };
        auto accum = [&](int c, float dv) {             // c = column of X, dv = dL/dX[c]
          if (REF && c == 0) { dx0 = dv; return; }
          if (c < C0 + APP_DIM) { df[c - C0] += dv; return; }
          if (c < PF) { if (REF) dr[c - C0 - APP_DIM] += dv; return; }          // view direction: no parameters (VM)
          if (c < PF + NF) {
            const int o = (c - PF) >> 1, q = (c - PF) & 1;
            df[o] = fmaf(dv * (q ? 2.0f : 1.0f), xval(c + NF), df[o]);
          } else if (c < PV) {
            const int o = (c - PF - NF) >> 1, q = (c - PF - NF) & 1;
            df[o] = fmaf(-dv * (q ? 2.0f : 1.0f), xval(c - NF), df[o]);
          } else if (REF && c < PV + NV) {
            const int o = (c - PV) >> 1, q = (c - PV) & 1;
            dr[o] = fmaf(dv * (q ? 2.0f : 1.0f), xval(c + NV), dr[o]);
          } else if (REF && c < PV + 2 * NV) {
            const int o = (c - PV - NV) >> 1, q = (c - PV - NV) & 1;
            dr[o] = fmaf(-dv * (q ? 2.0f : 1.0f), xval(c - NV), dr[o]);
          }
        };
#pragma unroll
        for (int cb = 0; cb < 5; ++cb) {
          float d[32];
          tmem_ld32(lane_addr + cWork + cb * 32, d);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = cb * 32 + j;
            if (c < IN_C) accum(c, d[j]);
          }
        }
        if (REF) {
          // x[0] = -d.n; reflection = 2 (d.n) n - d; penalty = sum w relu(-d.n)^2; n = v / |v|
          float ddot = -dx0 + 2.0f * (nh[0] * dr[0] + nh[1] * dr[1] + nh[2] * dr[2]);
          ddot -= dpen * wgt * 2.0f * fmaxf(-dotp, 0.0f);
          float dn[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) dn[c] = 2.0f * dotp * dr[c] + ddot * vd[c];
          const float proj = nh[0] * dn[0] + nh[1] * dn[1] + nh[2] * dn[2];
          const bool live = nraw2 > 1e-30f;          // degenerate normal: jt.normalize clamps, no gradient (padding rows: all zero)
#pragma unroll
          for (int c = 0; c < 3; ++c) df[APP_DIM + c] = live ? (dn[c] - nh[c] * proj) * ninv : 0.0f;
          df[APP_DIM + 3] = gw[0]; df[APP_DIM + 4] = gw[1]; df[APP_DIM + 5] = gw[2];     // d rgb_d = w g
          df[APP_DIM + 6] = tint > 0.0f ? dtint : 0.0f;                                  // tint = relu(specular_linear(h))
#pragma unroll
          for (int j = 0; j < 8; ++j) dhb[j] += df[APP_DIM + j];
        }
        // dF -> X columns 0..NH (bf16), in place
#pragma unroll
        for (int g = 0; g < NH / 8; ++g)
          *reinterpret_cast<uint4*>(xrow + g * kLbo) =
              make_uint4(pack_bf16(df[g * 8 + 0], df[g * 8 + 1]), pack_bf16(df[g * 8 + 2], df[g * 8 + 3]),
                         pack_bf16(df[g * 8 + 4], df[g * 8 + 5]), pack_bf16(df[g * 8 + 6], df[g * 8 + 7]));
      }
      publish();
      // ---- dBasis^T = H^T dF (two windows: channels 0..127 and 16..143), flushed per tile ------------------------------------------
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cDBt1, smem_desc(aH + s * 256, 128, kLbo), smem_desc(aX + s * 256, 128, kLbo),
                    instr_desc(128, NH) | MN, s > 0);
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16(tmem + cDBt2, smem_desc(aH + 2 * kLbo + s * 256, 128, kLbo), smem_desc(aX + s * 256, 128, kLbo),
                    instr_desc(128, NH) | MN, s > 0);
        umma_commit(mma_bar);
      }
      mma_wait();
      {
        // TMEM lane = appearance channel h, columns = stacked head outputs
        float v[32];
#pragma unroll
        for (int cb = 0; cb < NH; cb += 16) {
          tmem_ld16(lane_addr + cDBt1 + cb, v);
          float* gb = Bp.g.basis_t + (size_t)row * NH + cb;
#pragma unroll
          for (int i = 0; i < 16; i += 4) red_add_v4(gb + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        if (warp == 3) {             // second window: lane l = channel 16 + l; only channels 128..143 (rows >= 112) are new
#pragma unroll
          for (int cb = 0; cb < NH; cb += 16) {
            tmem_ld16(lane_addr + cDBt2 + cb, v);        // .sync.aligned: the whole warp executes the load
            if (row >= 112) {
              float* gb = Bp.g.basis_t + (size_t)(16 + row) * NH + cb;
#pragma unroll
              for (int i = 0; i < 16; i += 4) red_add_v4(gb + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          }
        }
      }
      fence_before();
      row_sync();                  // every lane has read the dBasis^T windows before dH overwrites those columns
      // ---- dH = dF basis ------------------------------------------------------------------------------------------------------------
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int s = 0; s < NH / 16; ++s)
          umma_bf16(tmem + cWork, smem_desc(aX + s * 2 * kLbo, kLbo, 128), smem_desc(aB0 + s * 256, 128, NH * 16),
                    instr_desc(128, K0) | kIdescBMajorMN, s > 0);
        umma_commit(mma_bar);
      }
      mma_wait();
      {
        // ---- dH -> shared memory (bf16, over H) for the scatter ----------------------------------------------------------------------
        float v[32];
        uint8_t* hrow = sH + row * 16;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          tmem_ld32(lane_addr + cWork + cb * 32, v);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(hrow + (cb * 4 + g) * kLbo) =
                make_uint4(pack_bf16(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16(v[g * 8 + 2], v[g * 8 + 3]),
                           pack_bf16(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16(v[g * 8 + 6], v[g * 8 + 7]));
        }
        float t16[16];
        tmem_ld16(lane_addr + cWork + 128, t16);
#pragma unroll
        for (int g = 0; g < 2; ++g)
          *reinterpret_cast<uint4*>(hrow + (16 + g) * kLbo) =
              make_uint4(pack_bf16(t16[g * 8 + 0], t16[g * 8 + 1]), pack_bf16(t16[g * 8 + 2], t16[g * 8 + 3]),
                         pack_bf16(t16[g * 8 + 4], t16[g * 8 + 5]), pack_bf16(t16[g * 8 + 6], t16[g * 8 + 7]));
      }
      fence_before();
    }
    first = false;
    __syncthreads();
    // ================================ bias column sums + scatter (all warps) ===========================================
    if (warp >= kRowWarps) {
      // warps 4-7: db2 += column sums of G2; warps 8-11: db1 += column sums of G1 (both still intact in shared memory)
      const int c = (tid - kRowWarps * 32) & 127;
      const uint8_t* base = (warp < kRowWarps + 4 ? sG2 : sY1) + (c >> 3) * kLbo + (c & 7) * 2;
      float a = 0.0f;
#pragma unroll 8
      for (int r = 0; r < kRows; ++r) a += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(base + r * 16));
      dbias += a;
    }
    {
      const int q = lane & 3;
#pragma unroll 1
      for (int grow = warp * 8 + (lane >> 2); grow < kRows; grow += (kThreads / 32) * 8) {
        const uint32_t ge = tile_base + grow;
        if (ge >= n_ent) continue;
        const float4 uw = __ldg(P.ws.ent_u + ge);
        const float u[3] = {uw.x, uw.y, uw.z};
        Axis ax[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) ax[i] = axis_taps(u[i], m.grid[i]);
        const uint8_t* hrow = sH + grow * 16;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const VmTaps t = vm_taps(m, ax, kk);
          float* gp = Bp.g.app_plane[kk];
          float* gl = Bp.g.app_line[kk];
#pragma unroll
          for (int c = q * 4; c < CA; c += 16) {
            float4 pv, lv;
            vm_sample4(m.app_plane[kk], m.app_line[kk], t, CA, c, pv, lv);
            const int k = kk * CA + c;
            const uint2 dh = *reinterpret_cast<const uint2*>(hrow + (k >> 3) * kLbo + (k & 7) * 2);
            const float dx = bf_lo(dh.x), dy = bf_hi(dh.x), dz = bf_lo(dh.y), dw = bf_hi(dh.y);
            const float px = dx * lv.x, py = dy * lv.y, pz = dz * lv.z, pw = dw * lv.w;   // d plane value
            const float lx = dx * pv.x, ly = dy * pv.y, lz = dz * pv.z, lw = dw * pv.w;   // d line value
            red_add_v4(gp + (size_t)t.o00 * CA + c, px * t.nw, py * t.nw, pz * t.nw, pw * t.nw);
            red_add_v4(gp + (size_t)t.o01 * CA + c, px * t.ne, py * t.ne, pz * t.ne, pw * t.ne);
            red_add_v4(gp + (size_t)t.o10 * CA + c, px * t.sw, py * t.sw, pz * t.sw, pw * t.sw);
            red_add_v4(gp + (size_t)t.o11 * CA + c, px * t.se, py * t.se, pz * t.se, pw * t.se);
            red_add_v4(gl + (size_t)t.l0 * CA + c, lx * t.lw0, ly * t.lw0, lz * t.lw0, lw * t.lw0);
            red_add_v4(gl + (size_t)t.l1 * CA + c, lx * t.lw1, ly * t.lw1, lz * t.lw1, lw * t.lw1);
          }
        }
      }
    }
    __syncthreads();
  }

  // ================================ flush of the persistent accumulators ==================================================
  if (!first) {
    if (warp < kRowWarps) {
      fence_after();
      const int o = tid;          // TMEM lane = output unit of the layer
      float v[32];
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        tmem_ld32(lane_addr + cDW2 + cb * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(Bp.g.w2_t + (size_t)(cb * 32 + i) * kFeatureC + o, v[i]);
      }
#pragma unroll 1
      for (int cb = 0; cb < 5; ++cb) {
        tmem_ld32(lane_addr + cDW1 + cb * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (cb * 32 + i < IN_C) atomicAdd(Bp.g.w1_t + (size_t)(cb * 32 + i) * kFeatureC + o, v[i]);
      }
      float t16[16];
      tmem_ld16(lane_addr + cDW3, t16);
#pragma unroll
      for (int j = 0; j < 3; ++j) atomicAdd(Bp.g.w3 + j * kFeatureC + o, t16[kDoCol - (kDoChunk - 1) * 8 + j]);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float s = warp_sum(db3[j]);
        if (lane == 0) atomicAdd(Bp.g.b3 + j, s);
      }
      if (REF) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float s = warp_sum(dhb[j]);
          if (lane == 0) atomicAdd(Bp.g.head_bias + APP_DIM + j, s);
        }
      }
    } else {
      const int c = (tid - kRowWarps * 32) & 127;
      atomicAdd((warp < kRowWarps + 4 ? Bp.g.b2 : Bp.g.b1) + c, dbias);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace bwdtc

int launch_app_bwd_tc(const BwdParams& B, int num_sms, cudaStream_t stream) {
  using namespace bwdtc;
  const TvmModel& m = B.f.m;
  TVM_REQUIRE(m.n_app == CA && m.app_dim == APP_DIM && m.fea_pe == 2 && m.view_pe == 2 &&
              m.feature_c == 128, "tensor-core appearance backward supports n_app=48, app_dim=27, fea_pe=view_pe=2, featureC=128");
  TVM_REQUIRE(m.tc_weights != nullptr || m.tc_weights_bwd != nullptr, "TvmModel.tc_weights is NULL: call tvm_pack_mlp_tc first");
  const bool ref = m.variant == TVM_VARIANT_REF;
  const tc::Image img(CA, in_mlp_c(m), head_ld(m));
  const uint32_t w_bytes = ref ? img.off_f32 : img.bytes;
  const size_t smem = ((w_bytes + 127) & ~127u) + (size_t)kRows * (K0 + K1 + 128 + 128) * 2 + 64;
  TVM_REQUIRE(smem <= 227 * 1024, "k_app_bwd_tc shared memory");
  auto kern = ref ? k_app_bwd_tc<true> : k_app_bwd_tc<false>;
  TVM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<num_sms, kThreads, smem, stream>>>(B);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvm
