// Per-ray / per-sample geometry of the NeRF++ background (nerfplusplus.py:196-237), shared by the fp32 and the
// tensor-core background kernels.
#pragma once
#include "tvm_common.cuh"

namespace tvm {

constexpr int kBgSamples = TVM_NPP_BG_SAMPLES;
constexpr int kPosDim = 20, kDirDim = 15, kBgHid = 64;

// Operand shapes and byte offsets of the background network's tensor-core weight image (tvm_pack_bg_tc), shared by the
// forward (tvm_bg_tc.cu) and the backward (tvm_bg_bwd_tc.cu) kernels.  Every B operand is a K-major core-matrix image
// (tvm_tc.cuh): element (n, k) at (k/8) * (N*16) + n*16 + (k%8)*2.
namespace bgimg {
constexpr int kPosK = 32;               // position block of the A operand (20 embedding columns, padded)
constexpr int kOneCol = 20;             // constant 1.0: carries the biases
constexpr int kAK = kPosK + kFeatureC;  // 160 columns
constexpr int kN3 = 80;                 // 64 hidden rgb units + sigma + padding (N % 16 == 0)
constexpr int kN4 = 16;                 // 3 colour channels, padded
constexpr int kK1 = kFeatureC + 16;     // hidden + the 16 position columns that hold the one-column
constexpr uint32_t kOffW0 = 0;                                   // [N 128][K  32]  base_layers.0 (+ b0 in row 20)
constexpr uint32_t kOffW1 = kOffW0 + kPosK * kFeatureC * 2;     // [N 128][K 144]  base_layers.1: hidden, then position columns 16..31 (b1)
constexpr uint32_t kOffW2 = kOffW1 + kK1 * kFeatureC * 2;       // [N 128][K 160]  base_layers.2: position block (b2 in row 20), then hidden
constexpr uint32_t kOffW3 = kOffW2 + kAK * kFeatureC * 2;       // [N  80][K 128]  folded colour layer | sigma head | 0
constexpr uint32_t kOffW4 = kOffW3 + kFeatureC * kN3 * 2;       // [N  16][K  64]  rgb_layers.2
constexpr uint32_t kOffF32 = kOffW4 + kBgHid * kN4 * 2;         // fp32 tail: b_sigma, b_rgb[3]
constexpr uint32_t kImageBytes = kOffF32 + 16;
}  // namespace bgimg

struct BgRay {
  float p_sphere[3], axis[3], cross_ap[3];   // point on the sphere, rotation axis, axis x p_sphere
  float axis_dot;                            // axis . p_sphere
  float pmn, phi;                            // |p_mid|, asin(|p_mid| / R)
};

// per-ray part of depth2pts_outside (nerfplusplus.py:212-225)
__device__ __forceinline__ void bg_ray_setup(const float* ray6, float R, BgRay& g) {
  const float o[3] = {ray6[0], ray6[1], ray6[2]}, d[3] = {ray6[3], ray6[4], ray6[5]};
  const float dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const float d1 = -(d[0] * o[0] + d[1] * o[1] + d[2] * o[2]) / dd;
  float pm[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pm[i] = o[i] + d1 * d[i];
  g.pmn = sqrtf(pm[0] * pm[0] + pm[1] * pm[1] + pm[2] * pm[2]);
  const float cosd = 1.0f / sqrtf(dd);
  const float d2 = sqrtf(R * R - g.pmn * g.pmn) * cosd;
#pragma unroll
  for (int i = 0; i < 3; ++i) g.p_sphere[i] = o[i] + (d1 + d2) * d[i];
  float ax[3] = {o[1] * g.p_sphere[2] - o[2] * g.p_sphere[1], o[2] * g.p_sphere[0] - o[0] * g.p_sphere[2],
                 o[0] * g.p_sphere[1] - o[1] * g.p_sphere[0]};
  const float inv = 1.0f / sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) g.axis[i] = ax[i] * inv;
  g.cross_ap[0] = g.axis[1] * g.p_sphere[2] - g.axis[2] * g.p_sphere[1];
  g.cross_ap[1] = g.axis[2] * g.p_sphere[0] - g.axis[0] * g.p_sphere[2];
  g.cross_ap[2] = g.axis[0] * g.p_sphere[1] - g.axis[1] * g.p_sphere[0];
  g.axis_dot = g.axis[0] * g.p_sphere[0] + g.axis[1] * g.p_sphere[1] + g.axis[2] * g.p_sphere[2];
  g.phi = asinf(g.pmn / R);
}

// torch.linspace(0, R, 512)[i] (symmetric evaluation) and perturb_samples (nerfplusplus.py:196-205)
__device__ __forceinline__ float bg_lin(int i, float R) {
  const float step = R / (float)(kBgSamples - 1);
  return i < kBgSamples / 2 ? (float)i * step : R - (float)(kBgSamples - 1 - i) * step;
}
__device__ __forceinline__ float bg_depth(int i, float R, const float* rnd) {
  const float f0 = bg_lin(i, R);
  float lower = f0, upper = f0;
  if (i > 0) lower = 0.5f * (f0 + bg_lin(i - 1, R));
  if (i < kBgSamples - 1) upper = 0.5f * (bg_lin(i + 1, R) + f0);
  return lower + (upper - lower) * rnd[i];
}

}  // namespace tvm
