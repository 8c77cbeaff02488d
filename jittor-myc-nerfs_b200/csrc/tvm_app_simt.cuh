// fp32 appearance head building blocks shared by the forward (k_app_simt) and backward (k_app_bwd)
// kernels: one CTA of 256 threads works on a tile of 64 entries; activations live in shared memory as
// [64][st] fp32 tiles with st == 4 (mod 32) so that both access patterns are bank-conflict free:
//   thread-per-row float4 reads (8 lanes x 16 B per wavefront land in 32 distinct banks) and
//   row-broadcast / lane-per-column reads used by the weight-gradient products.
#pragma once
#include "tvm_common.cuh"

namespace tvm {

constexpr int kAppTile = 64;
constexpr int kAppThreads = 256;

__host__ __device__ inline int app_tile_stride(int n_app, int in_c) {
  int k = 3 * n_app;
  if (in_c > k) k = in_c;
  if (kFeatureC > k) k = kFeatureC;
  return (k + 31) / 32 * 32 + 4;
}

__device__ __forceinline__ void fma4(float* acc, float x, const float4 w) {
  acc[0] = fmaf(x, w.x, acc[0]);
  acc[1] = fmaf(x, w.y, acc[1]);
  acc[2] = fmaf(x, w.z, acc[2]);
  acc[3] = fmaf(x, w.w, acc[3]);
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Gathers the 3*Ca appearance product vector of one tile into H and the view directions into
// X[.., app_dim .. app_dim+3).  8 warps x 8 entries, 4 lanes per entry (tensoRF.py:228-243).
__device__ __forceinline__ void app_gather_tile(const FwdParams& P, uint32_t tile_base, uint32_t n_ent,
                                                float* H, float* X, int st) {
  const TvmModel& m = P.m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ca = m.n_app;
  const int row = warp * 8 + (lane >> 2), q = lane & 3;
  const uint32_t e = tile_base + row;
  float* h = H + row * st;
  if (e < n_ent) {
    const uint2 en = P.ws.ent[e];
    float u[3], dir[3];
    entry_coords(m, P.rays, P.jitter, en.x, en.y, P.S, u, dir);
    Axis ax[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) ax[i] = axis_taps(u[i], m.grid[i]);
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      const VmTaps t = vm_taps(m, ax, kk);
      for (int c = q * 4; c < Ca; c += 16) {
        float4 pv, lv;
        vm_sample4(m.app_plane[kk], m.app_line[kk], t, Ca, c, pv, lv);
        *reinterpret_cast<float4*>(h + kk * Ca + c) = make_float4(pv.x * lv.x, pv.y * lv.y, pv.z * lv.z, pv.w * lv.w);
      }
    }
    if (q < 3) X[row * st + col_dir(m) + q] = dir[q];
  } else {
    for (int c = q * 4; c < 3 * Ca; c += 16) *reinterpret_cast<float4*>(h + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < 3) X[row * st + col_dir(m) + q] = 0.0f;
  }
}

// basis_mat (+ REFTensoRF heads) and positional encoding (tensoRF.py:244; tensorBase.py:9-15,76-83;
// REFTensoRF.py:107-133,216-232).  Thread (row, part) owns head outputs [part*NH/4, (part+1)*NH/4);
// NH = 32 (basis only) or 48 (basis | normal | diffuse | specular | rho).  HD [64][8] is a side buffer:
// in: raw head outputs of the REF variant; out: {rgb_d[3], tint, .., ..} for the final colour.
template <int NH>
__device__ __forceinline__ void app_basis_pe(const FwdParams& P, const float* H, float* X, float* HD, int st,
                                             uint32_t tile_base, uint32_t n_ent) {
  constexpr int NP = NH / 4;
  const TvmModel& m = P.m;
  const bool ref = NH == TVM_REF_HEAD_LD;
  const int row = threadIdx.x & (kAppTile - 1), part = threadIdx.x >> 6;
  const int K = 3 * m.n_app;
  float acc[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) acc[i] = 0.0f;
  const float* h = H + row * st;
  const float* bt = m.basis_t + part * NP;
  for (int j = 0; j < K; j += 4) {
    const float4 x = lds4(h + j);
    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int i = 0; i < NP; i += 4) fma4(acc + i, xs[jj], ldg4(bt + (j + jj) * NH + i));
  }
  float* xr = X + row * st;
  const int c_feat = col_feat(m), c_dir = col_dir(m);
  const int pe_f = c_dir + 3;                            // start of sin(PE(features))
  const int pe_v = pe_f + 2 * m.fea_pe * m.app_dim;      // start of sin(PE(direction))
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int o = part * NP + i;
    if (o < m.app_dim) {
      const float f = acc[i];
      xr[c_feat + o] = f;
      float fr = 1.0f;
      for (int q = 0; q < m.fea_pe; ++q, fr *= 2.0f) {
        const float s = f * fr;
        xr[pe_f + o * m.fea_pe + q] = sinf(s);
        xr[pe_f + m.fea_pe * m.app_dim + o * m.fea_pe + q] = cosf(s);
      }
    } else if (ref && o < m.app_dim + 8) {
      HD[row * 8 + (o - m.app_dim)] = acc[i] + __ldg(m.head_bias + o);
    }
  }
  if (ref) __syncthreads();
  if (part == 3) {
    float d[3] = {xr[c_dir], xr[c_dir + 1], xr[c_dir + 2]};     // view direction left there by the gather
    if (ref) {
      // normal = normalize(normal_linear(h)); d = -view; dot = d.n; reflection = 2 dot n - d
      float* hd = HD + row * 8;
      float nx = hd[0], ny = hd[1], nz = hd[2];
      const float inv = 1.0f / sqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-30f));
      nx *= inv; ny *= inv; nz *= inv;
      const float dx = -d[0], dy = -d[1], dz = -d[2];
      const float dot = dx * nx + dy * ny + dz * nz;
      d[0] = 2.0f * dot * nx - dx;
      d[1] = 2.0f * dot * ny - dy;
      d[2] = 2.0f * dot * nz - dz;
      xr[0] = -dot;
      xr[c_dir] = d[0]; xr[c_dir + 1] = d[1]; xr[c_dir + 2] = d[2];
      const float r0 = hd[3], r1 = hd[4], r2 = hd[5], tint = fmaxf(hd[6], 0.0f);
      const float v0 = hd[0], v1 = hd[1], v2 = hd[2];
      hd[0] = r0; hd[1] = r1; hd[2] = r2; hd[3] = tint;
      hd[4] = v0; hd[5] = v1; hd[6] = v2; hd[7] = dot;       // raw normal and d.n for the backward pass
      const uint32_t e = tile_base + row;
      if (e < n_ent) {
        const float pe = fmaxf(-dot, 0.0f);
        P.ws.ent_pen[e] = pe * pe;
      }
      if (P.aux.penalty) {
        // part == 3 is two whole warps: one atomic per warp instead of one per entry (single-address contention)
        const float pen = fmaxf(-dot, 0.0f);
        const float v = warp_sum(e < n_ent ? P.ws.ent_w[e] * pen * pen : 0.0f);
        if ((threadIdx.x & 31) == 0 && v != 0.0f) atomicAdd(P.aux.penalty, v);
      }
    }
    for (int c = 0; c < 3; ++c) {
      float fr = 1.0f;
      for (int q = 0; q < m.view_pe; ++q, fr *= 2.0f) {
        const float s = d[c] * fr;
        xr[pe_v + c * m.view_pe + q] = sinf(s);
        xr[pe_v + 3 * m.view_pe + c * m.view_pe + q] = cosf(s);
      }
    }
    for (int c = P.in_mlp_c; c < ((P.in_mlp_c + 3) & ~3); ++c) xr[c] = 0.0f;
  }
}

// y[row][part*32 .. +32) = act(x[row][0..K) @ Wt[K][128] + b): thread (row, part); RELU or identity
template <bool RELU>
__device__ __forceinline__ void app_dense(const float* __restrict__ Wt, const float* __restrict__ bias,
                                          const float* xin, int K, float* yout, int st) {
  const int row = threadIdx.x & (kAppTile - 1), part = threadIdx.x >> 6;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 b = ldg4(bias + part * 32 + i);
    acc[i] = b.x; acc[i + 1] = b.y; acc[i + 2] = b.z; acc[i + 3] = b.w;
  }
  const float* x = xin + row * st;
  const float* w = Wt + part * 32;
  const int K4 = K & ~3;
  for (int j = 0; j < K4; j += 4) {
    const float4 xv = lds4(x + j);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int i = 0; i < 32; i += 4) fma4(acc + i, xs[jj], ldg4(w + (size_t)(j + jj) * kFeatureC + i));
  }
  for (int j = K4; j < K; ++j) {
    const float xv = x[j];
#pragma unroll
    for (int i = 0; i < 32; i += 4) fma4(acc + i, xv, ldg4(w + (size_t)j * kFeatureC + i));
  }
  float* y = yout + row * st + part * 32;
#pragma unroll
  for (int i = 0; i < 32; i += 4)
    *reinterpret_cast<float4*>(y + i) =
        RELU ? make_float4(fmaxf(acc[i], 0.f), fmaxf(acc[i + 1], 0.f), fmaxf(acc[i + 2], 0.f), fmaxf(acc[i + 3], 0.f))
             : make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
}

// pre-sigmoid output o (< 3) of the last layer for this thread's row
__device__ __forceinline__ float app_out_logit(const TvmModel& m, const float* Y2, int st, int row, int o) {
  float a = m.b3[o];
  const float* x = Y2 + row * st;
  const float* w = m.w3 + o * kFeatureC;
  for (int j = 0; j < kFeatureC; j += 4) {
    const float4 xv = lds4(x + j);
    const float4 wv = ldg4(w + j);
    a = fmaf(xv.x, wv.x, a);
    a = fmaf(xv.y, wv.y, a);
    a = fmaf(xv.z, wv.z, a);
    a = fmaf(xv.w, wv.w, a);
  }
  return a;
}

}  // namespace tvm
