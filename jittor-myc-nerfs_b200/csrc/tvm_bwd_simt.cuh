// Tile-level building blocks of the fp32 (CUDA-core) backward kernels: weight-gradient products of a 64-row tile,
// bias gradients and the data gradient of one dense layer.  Shared by k_app_bwd (tvm_backward.cu) and k_bg_bwd
// (tvm_bg_bwd.cu).
#pragma once
#include "tvm_app_simt.cuh"
#include "tvm_bwd.cuh"

namespace tvm {

// out[K][ldo] += A[64][0..K)^T . Bm[64][0..128)   (weight gradient of one dense layer for this tile)
__device__ __forceinline__ void wgrad_tile_128(const float* A, const float* Bm, int st, int K, float* out, int ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o4 = lane * 4;
  for (int j0 = warp * 4; j0 < K; j0 += 32) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
#pragma unroll 4
    for (int row = 0; row < kAppTile; ++row) {
      const float4 bv = lds4(Bm + row * st + o4);
      const float4 av = lds4(A + row * st + j0);     // broadcast; columns >= K are zero padding
      fma4(acc[0], av.x, bv);
      fma4(acc[1], av.y, bv);
      fma4(acc[2], av.z, bv);
      fma4(acc[3], av.w, bv);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (j0 + a < K) red_add_v4(out + (size_t)(j0 + a) * ldo + o4, acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
  }
}
// out[K][NH] += A[64][0..K)^T . Bm[64][0..NH): one thread per j
template <int NH>
__device__ __forceinline__ void wgrad_tile_heads(const float* A, const float* Bm, int st, int K, float* out) {
  const int j = threadIdx.x;
  if (j >= K) return;
  float acc[NH];
#pragma unroll
  for (int i = 0; i < NH; ++i) acc[i] = 0.0f;
  for (int row = 0; row < kAppTile; ++row) {
    const float a = A[row * st + j];
#pragma unroll
    for (int i = 0; i < NH; i += 4) fma4(acc + i, a, lds4(Bm + row * st + i));
  }
#pragma unroll
  for (int i = 0; i < NH; i += 4) red_add_v4(out + (size_t)j * NH + i, acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
}
// bias gradient: out[o] += sum_rows Bm[row][o], o < 128
__device__ __forceinline__ void bgrad_tile(const float* Bm, int st, float* out) {
  const int o = threadIdx.x;
  if (o >= kFeatureC) return;
  float a = 0.0f;
  for (int row = 0; row < kAppTile; ++row) a += Bm[row * st + o];
  atomicAdd(out + o, a);
}
// dIn[row][j] = sum_o dz[row][o] * Wt[j][o] for j in [0,K); optional ReLU mask from act (may alias dst).
template <bool MASK>
__device__ __forceinline__ void app_dense_bwd(const float* __restrict__ Wt, const float* dz, int K, float* dst,
                                              const float* act, int st) {
  const int row = threadIdx.x & (kAppTile - 1), part = threadIdx.x >> 6;
  const int JP = ((K + 3) / 4 + 7) / 8 * 8;        // columns per part, multiple of 8
  const int jbeg = part * JP, jend = min(K, jbeg + JP);
  const float* d = dz + row * st;
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    for (int o = 0; o < kFeatureC; o += 4) {
      const float4 dv = lds4(d + o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (j0 + i < jend) {   // warp-uniform
          const float4 w = ldg4(Wt + (size_t)(j0 + i) * kFeatureC + o);
          acc[i] = fmaf(dv.x, w.x, fmaf(dv.y, w.y, fmaf(dv.z, w.z, fmaf(dv.w, w.w, acc[i]))));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (j0 + i < jend) {
        float v = acc[i];
        if (MASK) v = act[row * st + j0 + i] > 0.0f ? v : 0.0f;
        dst[row * st + j0 + i] = v;
      }
    }
  }
}


}  // namespace tvm
