// Per-ray / per-sample geometry of the NeRF++ background (nerfplusplus.py:196-237), shared by the fp32 and the
// tensor-core background kernels.
#pragma once
#include "tvm_common.cuh"

namespace tvm {

constexpr int kBgSamples = TVM_NPP_BG_SAMPLES;
constexpr int kPosDim = 20, kDirDim = 15, kBgHid = 64;

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
