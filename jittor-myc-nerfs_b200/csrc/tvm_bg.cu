// NeRF++ background of tensorf-myc's NerfPlusPlus (models/nerfplusplus.py:272-318): for every ray whose
// foreground left bg_lambda > 0.1, 512 inverse-depth samples on the inverted sphere
// (depth2pts_outside :207-237), Embedder (:7-56) and MLPNet (:66-140), composited front to back;
// rgb_map += bg_lambda * bg_rgb_map.
//
//   k_bg_simt   fp32 path: one CTA per active ray, 8 tiles of 64 samples through the shared fp32 dense
//               building blocks (tvm_app_simt.cuh); the transmittance is carried across the tiles and
//               the ray stops once it falls below 1e-6.
//   k_bg_fold   weights-only product that folds base_remap (no activation) into rgb_layers.0.
#include "tvm_app_simt.cuh"
#include "tvm_bg.cuh"

namespace tvm {

// y[row][part*16 .. +16) = relu(x[row][0..128) @ Wt[128][64] + vbias): thread (row, part)
__device__ __forceinline__ void bg_hidden(const float* __restrict__ Wt, const float* vbias, const float* xin,
                                          float* yout, int st) {
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

// REFRESH: only re-derive ws.bg_rgb in fp32 (tvm_backward_npp after a tensor-core forward: the backward's suffix sums
// must be consistent with its own fp32 recompute); rgb_map, counters and aux are left alone.
template <bool REFRESH>
__global__ void __launch_bounds__(kAppThreads) k_bg_simt(const FwdParams P) {
  extern __shared__ __align__(16) float smem[];
  const int st = P.st;
  float* A = smem;                       // [64][st]: position embedding (cols 0..19) | layer-1 output (20..147); later hidden 64
  float* Bf = smem + kAppTile * st;      // [64][st]: layer-0 output, later layer-2 output
  float* SR = smem + 2 * kAppTile * st;  // [64][4] sigma, r, g, b
  float* DZ = SR + kAppTile * 4;         // [64] distance to the next sample
  float* VB = DZ + kAppTile;             // [64] per-ray bias of the hidden rgb layer
  float* FLAG = VB + kBgHid;             // [1] transmittance carried across tiles (written by warp 0)
  const TvmBgNet& bg = P.bg;
  const float R = P.m.radii;
  const int tid = threadIdx.x, row = tid & (kAppTile - 1), part = tid >> 6, lane = tid & 31;
  const uint32_t n_active = P.ws.n_entries[1];

  for (uint32_t idx = blockIdx.x; idx < n_active; idx += gridDim.x) {
    const uint32_t ray = P.ws.bg_list[idx];
    const float* ray6 = P.rays + 6 * (size_t)ray;
    const float* rnd = P.bg_rand + (size_t)ray * kBgSamples;
    BgRay g;
    bg_ray_setup(ray6, R, g);
    if (tid < kBgHid) {
      // view-direction embedding (Embedder, 2 freqs) through its slice of rgb_layers.0, plus the folded bias
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
      for (int j = 0; j < kDirDim; ++j) a = fmaf(e[j], bg.wv_t[j * kBgHid + tid], a);
      VB[tid] = a;
    }
    if (tid == 0) FLAG[0] = 1.0f;
    float T = 1.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;     // warp 0 only
    int tiles_done = 0;
    __syncthreads();

    for (int tile = 0; tile < kBgSamples / kAppTile; ++tile) {
      if (FLAG[0] < 1e-6f) break;                        // uniform: remaining weight mass < 1e-6
      ++tiles_done;
      // ---- geometry + embedding: flipped order j (near the sphere first), original index i = 511 - j
      if (part == 0) {
        const int j = tile * kAppTile + row, i = kBgSamples - 1 - j;
        const float z = bg_depth(i, R, rnd);
        DZ[row] = (i > 0) ? z - bg_depth(i - 1, R, rnd) : 1e10f;          // bg_dists, HUGE_NUMBER last (:299-300)
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
      app_dense<true>(bg.w0_t, bg.b0, A, kPosDim, Bf, st);
      __syncthreads();
      app_dense<true>(bg.w1_t, bg.b1, Bf, kFeatureC, A + kPosDim, st);     // skip connection: [input_pts | base]
      __syncthreads();
      app_dense<true>(bg.w2_t, bg.b2, A, kPosDim + kFeatureC, Bf, st);
      __syncthreads();
      if (part == 3) {                                                     // sigma = |w . base + b| (:128-129)
        float a = bg.b_sigma[0];
        const float* x = Bf + row * st;
        for (int j = 0; j < kFeatureC; j += 4) {
          const float4 xv = lds4(x + j), wv = ldg4(bg.w_sigma + j);
          a = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, a))));
        }
        SR[row * 4] = fabsf(a);
      }
      bg_hidden(bg.wf_t, VB, Bf, A, st);
      __syncthreads();
      if (part < 3) {
        float a = bg.b_rgb[part];
        const float* x = A + row * st;
        for (int j = 0; j < kBgHid; j += 4) {
          const float4 xv = lds4(x + j), wv = ldg4(bg.w_rgb + part * kBgHid + j);
          a = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, a))));
        }
        SR[row * 4 + 1 + part] = 1.0f / (1.0f + expf(-a));
      }
      __syncthreads();
      // ---- front-to-back compositing of the tile by warp 0: lane owns rows 2l, 2l+1 --------------
      if (tid < 32) {
        float al[2], vv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rr = 2 * lane + h;
          al[h] = 1.0f - expf(-SR[rr * 4] * DZ[rr]);
          vv[h] = 1.0f - al[h] + 1e-6f;                                     // TINY_NUMBER (:303)
        }
        float pref = vv[0] * vv[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float t = __shfl_up_sync(0xffffffffu, pref, o);
          if (lane >= o) pref *= t;
        }
        float excl = __shfl_up_sync(0xffffffffu, pref, 1);
        if (lane == 0) excl = 1.0f;
        const float T0 = T * excl, T1 = T0 * vv[0];
        const float w0 = al[0] * T0, w1 = al[1] * T1;
        const float* s0 = SR + (2 * lane) * 4;
        c0 += w0 * s0[1] + w1 * s0[5];
        c1 += w0 * s0[2] + w1 * s0[6];
        c2 += w0 * s0[3] + w1 * s0[7];
        T = T * __shfl_sync(0xffffffffu, pref, 31);
        if (lane == 0) FLAG[0] = T;
      }
      __syncthreads();
    }
    if (tid < 32) {
      c0 = warp_sum(c0);
      c1 = warp_sum(c1);
      c2 = warp_sum(c2);
      if (lane == 0 && REFRESH) {
        P.ws.bg_rgb[(size_t)ray * 3 + 0] = c0;
        P.ws.bg_rgb[(size_t)ray * 3 + 1] = c1;
        P.ws.bg_rgb[(size_t)ray * 3 + 2] = c2;
      } else if (lane == 0) {
        if (P.counters) {
          atomicAdd(&P.counters[TVM_CNT_BG_RAYS], 1ull);
          atomicAdd(&P.counters[TVM_CNT_BG_SAMPLES], (unsigned long long)(tiles_done * kAppTile));
        }
        const float lam = P.ws.bg_lambda[ray];
        P.rgb_map[(size_t)ray * 3 + 0] += lam * c0;       // nerfplusplus.py:314-317 (no clamp after the sum)
        P.rgb_map[(size_t)ray * 3 + 1] += lam * c1;
        P.rgb_map[(size_t)ray * 3 + 2] += lam * c2;
        P.ws.bg_rgb[(size_t)ray * 3 + 0] = c0;      // kept for tvm_backward_npp
        P.ws.bg_rgb[(size_t)ray * 3 + 1] = c1;
        P.ws.bg_rgb[(size_t)ray * 3 + 2] = c2;
        if (P.aux.bg_rgb_map) {
          P.aux.bg_rgb_map[(size_t)ray * 3 + 0] = c0;
          P.aux.bg_rgb_map[(size_t)ray * 3 + 1] = c1;
          P.aux.bg_rgb_map[(size_t)ray * 3 + 2] = c2;
        }
      }
    }
    __syncthreads();
  }
}

// wf_t[j][o] = sum_r rgb0_w[o][r] * remap_w[r][j];  bf[o] = sum_r rgb0_w[o][r] * remap_b[r] + rgb0_b[o];
// wv_t[v][o] = rgb0_w[o][256 + v]
__global__ void k_bg_fold(const float* __restrict__ remap_w, const float* __restrict__ remap_b,
                          const float* __restrict__ rgb0_w, const float* __restrict__ rgb0_b, float* __restrict__ wf_t,
                          float* __restrict__ bf, float* __restrict__ wv_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int IN = 256 + kDirDim;
  if (i < kFeatureC * kBgHid) {
    const int j = i / kBgHid, o = i % kBgHid;
    float a = 0.0f;
    for (int r = 0; r < 256; ++r) a = fmaf(rgb0_w[o * IN + r], remap_w[r * kFeatureC + j], a);
    wf_t[i] = a;
  } else if (i < kFeatureC * kBgHid + kBgHid) {
    const int o = i - kFeatureC * kBgHid;
    float a = rgb0_b[o];
    for (int r = 0; r < 256; ++r) a = fmaf(rgb0_w[o * IN + r], remap_b[r], a);
    bf[o] = a;
  } else if (i < kFeatureC * kBgHid + kBgHid + kDirDim * kBgHid) {
    const int t = i - kFeatureC * kBgHid - kBgHid, v = t / kBgHid, o = t % kBgHid;
    wv_t[t] = rgb0_w[o * IN + 256 + v];
  }
}

int launch_bg_tc(const FwdParams& P, int num_sms, cudaStream_t stream);   // tvm_bg_tc.cu

int launch_bg(const FwdParams& P, int num_sms, cudaStream_t stream) {
  const TvmBgNet& b = P.bg;
  TVM_REQUIRE(b.w0_t && b.b0 && b.w1_t && b.b1 && b.w2_t && b.b2 && b.w_sigma && b.b_sigma && b.wf_t && b.bf &&
              b.wv_t && b.w_rgb && b.b_rgb, "null TvmBgNet pointer");
  TVM_REQUIRE(P.m.radii > 0.0f, "radii must be positive");
  if (P.aux.bg_rgb_map) TVM_CHECK_CUDA(cudaMemsetAsync(P.aux.bg_rgb_map, 0, (size_t)P.n * 12, stream));
  if ((P.flags & TVM_MLP_MASK) != TVM_MLP_FP32) return launch_bg_tc(P, num_sms, stream);   // bf16 operands in both tensor-core modes
  const size_t smem = ((size_t)kAppTile * 2 * P.st + kAppTile * 4 + kAppTile + kBgHid + 4) * sizeof(float);
  TVM_CHECK_CUDA(cudaFuncSetAttribute(k_bg_simt<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bg_simt<false><<<num_sms * 2, kAppThreads, smem, stream>>>(P);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// fp32 re-evaluation of ws.bg_rgb for the rays of ws.bg_list (see k_bg_simt<REFRESH>)
int launch_bg_refresh(const FwdParams& P, int num_sms, cudaStream_t stream) {
  const size_t smem = ((size_t)kAppTile * 2 * P.st + kAppTile * 4 + kAppTile + kBgHid + 4) * sizeof(float);
  TVM_CHECK_CUDA(cudaFuncSetAttribute(k_bg_simt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bg_simt<true><<<num_sms * 2, kAppThreads, smem, stream>>>(P);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvm

using namespace tvm;

extern "C" int tvm_bg_fold(const float* remap_w, const float* remap_b, const float* rgb0_w, const float* rgb0_b,
                           float* wf_t, float* bf, float* wv_t, void* stream) {
  TVM_REQUIRE(remap_w && remap_b && rgb0_w && rgb0_b && wf_t && bf && wv_t, "bad arguments");
  const int n = kFeatureC * kBgHid + kBgHid + kDirDim * kBgHid;
  k_bg_fold<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(remap_w, remap_b, rgb0_w, rgb0_b, wf_t, bf, wv_t);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
