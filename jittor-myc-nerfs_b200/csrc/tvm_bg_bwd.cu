// Backward of the NeRF++ background (NerfPlusPlus.execute, models/nerfplusplus.py:283-317, under Jittor autograd):
// d rgb_map -> gradients of MLPNet's parameters (models/nerfplusplus.py:66-140).
//
//   k_bg_bwd        fp32, one CTA per evaluated ray (bg_lambda > 0.1), tiles of 64 samples in the forward kernel's order
//                   (k_bg_simt, tvm_bg.cu).  Per tile: recompute the network keeping every activation in shared memory,
//                   composite front to back (warp 0) to get w_j, T_j, then
//                     dL/d c_j     = g' w_j                      g' = bg_lambda * d rgb_map
//                     dL/d alpha_j = T_j g'.c_j - (sum_{i>j} w_i g'.c_i) / (1 - alpha_j + 1e-6)
//                   with the suffix sums as  g'.C_bg - prefix  (C_bg kept by the forward pass in the workspace), so ONE
//                   forward sweep suffices; back-propagate through sigmoid / |.| / the skip MLP with the tile-level
//                   products of tvm_bwd_simt.cuh and red.global.add the weight gradients.  Sample positions and view
//                   directions carry no parameters, so nothing flows into them.
//   k_bg_fold_bwd   transpose of k_bg_fold: gradients of the folded colour layer -> base_remap_layers.0, rgb_layers.0.
#include <stdlib.h>
#include "tvm_bwd_simt.cuh"
#include "tvm_bg.cuh"

namespace tvm {

struct BgBwdParams {
  FwdParams f;
  const float* d_rgb_map;
  TvmBgGrads g;
};

// yout[row][part*16 .. +16) = relu(xin[row][0..128) @ Wt[128][64] + vbias)   (the forward's bg_hidden, tvm_bg.cu)
__device__ __forceinline__ void bgb_hidden(const float* __restrict__ Wt, const float* vbias, const float* xin, float* yout,
                                           int st) {
  const int row = threadIdx.x & (kAppTile - 1), part = threadIdx.x >> 6;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = vbias[part * 16 + i];
  const float* x = xin + row * st;
  const float* w = Wt + part * 16;
  for (int j = 0; j < kFeatureC; j += 4) {
    const float4 xv = lds4(x + j);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int i = 0; i < 16; i += 4) fma4(acc + i, xs[jj], ldg4(w + (size_t)(j + jj) * kBgHid + i));
  }
  float* y = yout + row * st + part * 16;
#pragma unroll
  for (int i = 0; i < 16; i += 4)
    *reinterpret_cast<float4*>(y + i) = make_float4(fmaxf(acc[i], 0.f), fmaxf(acc[i + 1], 0.f), fmaxf(acc[i + 2], 0.f),
                                                    fmaxf(acc[i + 3], 0.f));
}

__global__ void __launch_bounds__(kAppThreads) k_bg_bwd(const BgBwdParams B) {
  extern __shared__ __align__(16) float smem[];
  const FwdParams& P = B.f;
  const int st = P.st;
  float* A = smem;                        // [64][st]: pos embedding (0..19) | layer-1 output (20..147) -> delta1 in place
  float* Y0 = smem + kAppTile * st;       // [64][st]: layer-0 output -> delta0 in place
  float* Y2 = smem + 2 * kAppTile * st;   // [64][st]: layer-2 output -> delta2 in place
  float* HID = smem + 3 * kAppTile * st;  // [64][st]: hidden colour layer (64) -> its delta in place
  float* SR = smem + 4 * kAppTile * st;   // [64][4]  sigma (signed pre-activation in the sign bit convention below), r, g, b
  float* D = SR + kAppTile * 4;           // [64][4]  d logit r, g, b, d sigma pre-activation
  float* DZ = D + kAppTile * 4;           // [64]     distance to the next sample
  float* SG = DZ + kAppTile;              // [64]     sign of the sigma pre-activation
  float* VB = SG + kAppTile;              // [64]     per-ray bias of the hidden colour layer
  float* DVB = VB + kBgHid;               // [64]     its gradient, summed over the ray
  float* FLAG = DVB + kBgHid;             // [4]      transmittance carried across tiles
  const TvmBgNet& bg = P.bg;
  const float R = P.m.radii;
  const int tid = threadIdx.x, row = tid & (kAppTile - 1), part = tid >> 6, lane = tid & 31;
  const uint32_t n_active = P.ws.n_entries[1];

  for (uint32_t idx = blockIdx.x; idx < n_active; idx += gridDim.x) {
    const uint32_t ray = P.ws.bg_list[idx];
    const float* ray6 = P.rays + 6 * (size_t)ray;
    const float* rnd = P.bg_rand + (size_t)ray * kBgSamples;
    const float lam = P.ws.bg_lambda[ray];
    const float g0 = lam * B.d_rgb_map[(size_t)ray * 3 + 0], g1 = lam * B.d_rgb_map[(size_t)ray * 3 + 1],
                g2 = lam * B.d_rgb_map[(size_t)ray * 3 + 2];
    if (g0 == 0.0f && g1 == 0.0f && g2 == 0.0f) continue;     // uniform over the CTA
    const float gtot = g0 * P.ws.bg_rgb[(size_t)ray * 3 + 0] + g1 * P.ws.bg_rgb[(size_t)ray * 3 + 1] +
                       g2 * P.ws.bg_rgb[(size_t)ray * 3 + 2];
    BgRay g;
    bg_ray_setup(ray6, R, g);
    float e[kDirDim];                     // view-direction embedding (threads < 64 use it twice)
    if (tid < kBgHid) {
      const float dn = 1.0f / sqrtf(ray6[3] * ray6[3] + ray6[4] * ray6[4] + ray6[5] * ray6[5]);
      const float v[3] = {ray6[3] * dn, ray6[4] * dn, ray6[5] * dn};
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
      for (int j = 0; j < kDirDim; ++j) a = fmaf(e[j], bg.wv_t[j * kBgHid + tid], a);
      VB[tid] = a;
      DVB[tid] = 0.0f;
    }
    if (tid == 0) FLAG[0] = 1.0f;
    float T = 1.0f, carry = 0.0f;         // warp 0 only
    __syncthreads();

    for (int tile = 0; tile < kBgSamples / kAppTile; ++tile) {
      if (FLAG[0] < 1e-6f) break;         // the forward stopped here as well
      // ---- forward recompute (k_bg_simt) --------------------------------------------------------
      if (part == 0) {
        const int j = tile * kAppTile + row, i = kBgSamples - 1 - j;
        const float z = bg_depth(i, R, rnd);
        DZ[row] = (i > 0) ? z - bg_depth(i - 1, R, rnd) : 1e10f;
        const float theta = asinf(g.pmn * z / (R * R));
        const float ang = g.phi - theta;
        float sa, ca;
        sincosf(ang, &sa, &ca);
        float x[4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          x[c] = g.p_sphere[c] * ca + g.cross_ap[c] * sa + g.axis[c] * g.axis_dot * (1.0f - ca);
        x[3] = z;
        float* a = A + row * st;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          a[c] = x[c];
          a[4 + c] = sinf(x[c]);
          a[8 + c] = cosf(x[c]);
          a[12 + c] = sinf(2.0f * x[c]);
          a[16 + c] = cosf(2.0f * x[c]);
        }
      }
      __syncthreads();
      app_dense<true>(bg.w0_t, bg.b0, A, kPosDim, Y0, st);
      __syncthreads();
      app_dense<true>(bg.w1_t, bg.b1, Y0, kFeatureC, A + kPosDim, st);
      __syncthreads();
      app_dense<true>(bg.w2_t, bg.b2, A, kPosDim + kFeatureC, Y2, st);
      __syncthreads();
      if (part == 3) {                                   // sigma = |w . base + b| (nerfplusplus.py:128-129)
        float a = bg.b_sigma[0];
        const float* x = Y2 + row * st;
        for (int j = 0; j < kFeatureC; j += 4) {
          const float4 xv = lds4(x + j), wv = ldg4(bg.w_sigma + j);
          a = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, a))));
        }
        SR[row * 4] = fabsf(a);
        SG[row] = a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f);
      }
      bgb_hidden(bg.wf_t, VB, Y2, HID, st);
      __syncthreads();
      if (part < 3) {
        float a = bg.b_rgb[part];
        const float* x = HID + row * st;
        for (int j = 0; j < kBgHid; j += 4) {
          const float4 xv = lds4(x + j), wv = ldg4(bg.w_rgb + part * kBgHid + j);
          a = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, a))));
        }
        SR[row * 4 + 1 + part] = 1.0f / (1.0f + expf(-a));
      }
      __syncthreads();
      // ---- compositing + its backward by warp 0: lane owns rows 2l, 2l+1 ---------------------------
      if (tid < 32) {
        float al[2], vv[2], gc[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rr = 2 * lane + h;
          al[h] = 1.0f - expf(-SR[rr * 4] * DZ[rr]);
          vv[h] = 1.0f - al[h] + 1e-6f;
          gc[h] = g0 * SR[rr * 4 + 1] + g1 * SR[rr * 4 + 2] + g2 * SR[rr * 4 + 3];     // g' . c_j
        }
        float pref = vv[0] * vv[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float t = __shfl_up_sync(0xffffffffu, pref, o);
          if (lane >= o) pref *= t;
        }
        float excl = __shfl_up_sync(0xffffffffu, pref, 1);
        if (lane == 0) excl = 1.0f;
        const float Tj[2] = {T * excl, T * excl * vv[0]};
        const float w[2] = {al[0] * Tj[0], al[1] * Tj[1]};
        // inclusive prefix of q_j = w_j g'.c_j in sample order
        float ps = w[0] * gc[0] + w[1] * gc[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float t = __shfl_up_sync(0xffffffffu, ps, o);
          if (lane >= o) ps += t;
        }
        const float incl1 = carry + ps, incl0 = incl1 - w[1] * gc[1];
        const float incl[2] = {incl0, incl1};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rr = 2 * lane + h;
          const float dalpha = Tj[h] * gc[h] - (gtot - incl[h]) / vv[h];
          const float dsig = dalpha * DZ[rr] * (1.0f - al[h]);          // alpha = 1 - exp(-sigma dist)
          const float s0 = SR[rr * 4 + 1], s1 = SR[rr * 4 + 2], s2 = SR[rr * 4 + 3];
          D[rr * 4 + 0] = g0 * w[h] * s0 * (1.0f - s0);
          D[rr * 4 + 1] = g1 * w[h] * s1 * (1.0f - s1);
          D[rr * 4 + 2] = g2 * w[h] * s2 * (1.0f - s2);
          D[rr * 4 + 3] = dsig * SG[rr];
        }
        carry += __shfl_sync(0xffffffffu, ps, 31);
        T = T * __shfl_sync(0xffffffffu, pref, 31);
        if (lane == 0) FLAG[0] = T;
      }
      __syncthreads();
      // ---- colour head: d w_rgb, d b_rgb, d w_sigma, d b_sigma ---------------------------------------
      if (tid < 3 * kBgHid) {
        const int c = tid / kBgHid, h = tid % kBgHid;
        float a = 0.0f;
        for (int r = 0; r < kAppTile; ++r) a = fmaf(D[r * 4 + c], HID[r * st + h], a);
        atomicAdd(B.g.w_rgb + c * kBgHid + h, a);
      } else if (tid < 3 * kBgHid + 4) {
        const int c = tid - 3 * kBgHid;
        float a = 0.0f;
        for (int r = 0; r < kAppTile; ++r) a += D[r * 4 + c];
        atomicAdd(c < 3 ? B.g.b_rgb + c : B.g.b_sigma, a);
      }
      if (tid >= kFeatureC) {
        const int j = tid - kFeatureC;
        float a = 0.0f;
        for (int r = 0; r < kAppTile; ++r) a = fmaf(D[r * 4 + 3], Y2[r * st + j], a);
        atomicAdd(B.g.w_sigma + j, a);
      }
      __syncthreads();
      // ---- hidden colour layer: delta (in place of HID) ----------------------------------------------------
      {
        const float d0 = D[row * 4 + 0], d1 = D[row * 4 + 1], d2 = D[row * 4 + 2];
        float* y = HID + row * st + part * 16;
        const float* w = bg.w_rgb + part * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 yv = lds4(y + i);
          const float4 w0 = ldg4(w + i), w1 = ldg4(w + kBgHid + i), w2 = ldg4(w + 2 * kBgHid + i);
          float4 o;
          o.x = yv.x > 0.0f ? fmaf(d0, w0.x, fmaf(d1, w1.x, d2 * w2.x)) : 0.0f;
          o.y = yv.y > 0.0f ? fmaf(d0, w0.y, fmaf(d1, w1.y, d2 * w2.y)) : 0.0f;
          o.z = yv.z > 0.0f ? fmaf(d0, w0.z, fmaf(d1, w1.z, d2 * w2.z)) : 0.0f;
          o.w = yv.w > 0.0f ? fmaf(d0, w0.w, fmaf(d1, w1.w, d2 * w2.w)) : 0.0f;
          *reinterpret_cast<float4*>(y + i) = o;
        }
      }
      __syncthreads();
      // d wf_t [128][64] += Y2^T delta_h; d VB += column sums of delta_h
      wgrad_tile_heads<kBgHid>(Y2, HID, st, kFeatureC, B.g.wf_t);
      if (tid >= kFeatureC && tid < kFeatureC + kBgHid) {
        const int h = tid - kFeatureC;
        float a = 0.0f;
        for (int r = 0; r < kAppTile; ++r) a += HID[r * st + h];
        DVB[h] += a;
      }
      __syncthreads();
      // ---- layer 2: delta2 = relu'(Y2) * (delta_h . wf^T + d sigma_pre * w_sigma), in place of Y2 ---------------
      {
        float dh[kBgHid];
#pragma unroll
        for (int i = 0; i < kBgHid; i += 4) {
          const float4 v = lds4(HID + row * st + i);
          dh[i] = v.x; dh[i + 1] = v.y; dh[i + 2] = v.z; dh[i + 3] = v.w;
        }
        const float ds = D[row * 4 + 3];
        float* y = Y2 + row * st;
        for (int j = part * 32; j < part * 32 + 32; ++j) {
          const float* wt = bg.wf_t + (size_t)j * kBgHid;
          float a = ds * bg.w_sigma[j];
#pragma unroll
          for (int i = 0; i < kBgHid; i += 4) {
            const float4 w = ldg4(wt + i);
            a = fmaf(dh[i], w.x, fmaf(dh[i + 1], w.y, fmaf(dh[i + 2], w.z, fmaf(dh[i + 3], w.w, a))));
          }
          y[j] = y[j] > 0.0f ? a : 0.0f;
        }
      }
      __syncthreads();
      wgrad_tile_128(A, Y2, st, kPosDim + kFeatureC, B.g.w2_t, kFeatureC);
      bgrad_tile(Y2, st, B.g.b2);
      __syncthreads();
      // ---- layer 1: delta1 over the skip block A[:, 20..147] -----------------------------------------------------
      app_dense_bwd<true>(bg.w2_t + (size_t)kPosDim * kFeatureC, Y2, kFeatureC, A + kPosDim, A + kPosDim, st);
      __syncthreads();
      wgrad_tile_128(Y0, A + kPosDim, st, kFeatureC, B.g.w1_t, kFeatureC);
      bgrad_tile(A + kPosDim, st, B.g.b1);
      __syncthreads();
      // ---- layer 0 ----------------------------------------------------------------------------------------------------
      app_dense_bwd<true>(bg.w1_t, A + kPosDim, kFeatureC, Y0, Y0, st);
      __syncthreads();
      wgrad_tile_128(A, Y0, st, kPosDim, B.g.w0_t, kFeatureC);
      bgrad_tile(Y0, st, B.g.b0);
      __syncthreads();
    }
    // VB = bf + wv^T e: d bf, d wv_t
    if (tid < kBgHid) {
      const float d = DVB[tid];
      atomicAdd(B.g.bf + tid, d);
#pragma unroll
      for (int j = 0; j < kDirDim; ++j) atomicAdd(B.g.wv_t + j * kBgHid + tid, e[j] * d);
    }
    __syncthreads();
  }
}

// Transpose of k_bg_fold (tvm_bg.cu): wf_t[j][o] = sum_r rgb0_w[o][r] remap_w[r][j], bf[o] = sum_r rgb0_w[o][r] remap_b[r]
// + rgb0_b[o], wv_t[v][o] = rgb0_w[o][256 + v]
__global__ void k_bg_fold_bwd(const float* __restrict__ remap_w, const float* __restrict__ remap_b,
                              const float* __restrict__ rgb0_w, const float* __restrict__ d_wf_t,
                              const float* __restrict__ d_bf, const float* __restrict__ d_wv_t, float* __restrict__ d_remap_w,
                              float* __restrict__ d_remap_b, float* __restrict__ d_rgb0_w, float* __restrict__ d_rgb0_b) {
  constexpr int IN = 256 + kDirDim;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = 256 * kFeatureC, n1 = n0 + 256, n2 = n1 + kBgHid * IN, n3 = n2 + kBgHid;
  if (i < n0) {                                   // d remap_w[r][j] = sum_o rgb0_w[o][r] d_wf_t[j][o]
    const int r = i / kFeatureC, j = i % kFeatureC;
    float a = 0.0f;
    for (int o = 0; o < kBgHid; ++o) a = fmaf(rgb0_w[o * IN + r], d_wf_t[j * kBgHid + o], a);
    d_remap_w[i] = a;
  } else if (i < n1) {                            // d remap_b[r] = sum_o rgb0_w[o][r] d_bf[o]
    const int r = i - n0;
    float a = 0.0f;
    for (int o = 0; o < kBgHid; ++o) a = fmaf(rgb0_w[o * IN + r], d_bf[o], a);
    d_remap_b[r] = a;
  } else if (i < n2) {
    const int t = i - n1, o = t / IN, r = t % IN;
    float a;
    if (r < 256) {                                // d rgb0_w[o][r] = sum_j d_wf_t[j][o] remap_w[r][j] + d_bf[o] remap_b[r]
      a = d_bf[o] * remap_b[r];
      for (int j = 0; j < kFeatureC; ++j) a = fmaf(d_wf_t[j * kBgHid + o], remap_w[r * kFeatureC + j], a);
    } else {
      a = d_wv_t[(r - 256) * kBgHid + o];
    }
    d_rgb0_w[t] = a;
  } else if (i < n3) {
    d_rgb0_b[i - n2] = d_bf[i - n2];
  }
}

int launch_bg_refresh(const FwdParams& P, int num_sms, cudaStream_t stream);   // tvm_bg.cu
int launch_bg_bwd_tc(const BwdParams& Bw, const TvmBgGrads& gr, int num_sms, cudaStream_t stream);   // tvm_bg_bwd_tc.cu

int launch_bg_bwd(const BwdParams& Bw, const TvmBgGrads& gr, int num_sms, cudaStream_t stream) {
  const TvmBgNet& b = Bw.f.bg;
  TVM_REQUIRE(b.w0_t && b.b0 && b.w1_t && b.b1 && b.w2_t && b.b2 && b.w_sigma && b.b_sigma && b.wf_t && b.bf &&
              b.wv_t && b.w_rgb && b.b_rgb, "null TvmBgNet pointer");
  TVM_REQUIRE(gr.w0_t && gr.b0 && gr.w1_t && gr.b1 && gr.w2_t && gr.b2 && gr.w_sigma && gr.b_sigma && gr.wf_t && gr.bf &&
              gr.wv_t && gr.w_rgb && gr.b_rgb, "null TvmBgGrads pointer");
  // tensor-core modes: the backward runs on the tensor cores as well (bf16 operands; its recompute is bit-identical to the
  // forward kernel, so the forward's bg_rgb is the sum it needs).  TVM_BG_BWD_FP32=1 keeps the fp32 kernel (A/B switch).
  static const bool force_fp32 = [] { const char* e = getenv("TVM_BG_BWD_FP32"); return e && e[0] == '1'; }();
  if ((Bw.f.flags & TVM_MLP_MASK) != TVM_MLP_FP32 && b.tc_weights != nullptr && !force_fp32)
    return launch_bg_bwd_tc(Bw, gr, num_sms, stream);
  // the forward ran on the tensor cores: its bg_rgb (bf16 network) is not the sum this kernel's fp32 recompute produces
  if ((Bw.f.flags & TVM_MLP_MASK) != TVM_MLP_FP32)
    if (int rc = launch_bg_refresh(Bw.f, num_sms, stream)) return rc;
  BgBwdParams B;
  B.f = Bw.f;
  B.d_rgb_map = Bw.d_rgb_map;
  B.g = gr;
  const size_t smem = ((size_t)kAppTile * 4 * B.f.st + kAppTile * (4 + 4 + 1 + 1) + 2 * kBgHid + 4) * sizeof(float);
  TVM_REQUIRE(smem <= 220 * 1024, "background backward tile does not fit shared memory");
  TVM_CHECK_CUDA(cudaFuncSetAttribute(k_bg_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bg_bwd<<<num_sms, kAppThreads, smem, stream>>>(B);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvm

using namespace tvm;

extern "C" int tvm_bg_fold_bwd(const float* remap_w, const float* remap_b, const float* rgb0_w, const float* d_wf_t,
                               const float* d_bf, const float* d_wv_t, float* d_remap_w, float* d_remap_b, float* d_rgb0_w,
                               float* d_rgb0_b, void* stream) {
  TVM_REQUIRE(remap_w && remap_b && rgb0_w && d_wf_t && d_bf && d_wv_t && d_remap_w && d_remap_b && d_rgb0_w && d_rgb0_b,
              "bad arguments");
  const int n = 256 * kFeatureC + 256 + kBgHid * (256 + kDirDim) + kBgHid;
  k_bg_fold_bwd<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(remap_w, remap_b, rgb0_w, d_wf_t, d_bf, d_wv_t, d_remap_w,
                                                                   d_remap_b, d_rgb0_w, d_rgb0_b);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
