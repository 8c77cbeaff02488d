// Per-sample arithmetic of the TensoRF-VM ray path, shared by every kernel.
//
// Everything that decides an INTEGER result (in-bbox mask, alpha-mask bit, tap indices) is written
// with explicitly rounded fp32 operations (no FMA contraction) in exactly the operation order of the
// reference python (tensorf-myc/models/tensorBase.py:340-360, 39-59, 223-224 and Jittor's
// grid_sample index arithmetic, assumption A1 of SURVEY.md §8c), so that mask bits and gathered
// texel indices are bit-identical to the oracle.  The functions are __host__ __device__ so that
// tests/host_emul can exercise the same source on the CPU-only build container.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/tvmrender.h"

#if defined(__CUDACC__)
#define TVM_HD __host__ __device__ __forceinline__
#else
#define TVM_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define TVM_ADD(a, b) __fadd_rn((a), (b))
#define TVM_SUB(a, b) __fsub_rn((a), (b))
#define TVM_MUL(a, b) __fmul_rn((a), (b))
#define TVM_DIV(a, b) __fdiv_rn((a), (b))
#else
// host build is compiled with -ffp-contract=off; plain operators are single IEEE roundings
#define TVM_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define TVM_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define TVM_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define TVM_DIV(a, b) ((float)((float)(a) / (float)(b)))
#endif

namespace tvm {

// matMode / vecMode of tensorBase.py:168-169: plane k spans axes (M0[k] -> W, M1[k] -> H), line k axis V[k]
TVM_HD int mat0(int k) { return k == 2 ? 1 : 0; }
TVM_HD int mat1(int k) { return k == 0 ? 1 : 2; }
TVM_HD int vecax(int k) { return 2 - k; }

struct RayMarch {
  float o[3], d[3];
  float t_min;   // entry distance clamped to [near, far]      (tensorBase.py:345-348); NeRF++: near
  float t_far;   // slab-test exit distance (uniform marching only; bounds the block loop, decides no mask bit)
  float jit;     // per-ray jitter (0 when !is_train)           (tensorBase.py:351-353)
  // TVM_SAMPLING_NPP (NerfPlusPlus.sample_ray, nerfplusplus.py:239-269)
  float step;          // (far - near) / (S - 1), far = exit of the radii-sphere
  const float* rnd;    // this ray's S draws of perturb_samples' U[0,1)
  int S;
};

// 3-term dot product in the reduction order of a sum over the last axis: (a0 b0 + a1 b1) + a2 b2
TVM_HD float dot3_seq(const float a[3], const float b[3]) {
  return TVM_ADD(TVM_ADD(TVM_MUL(a[0], b[0]), TVM_MUL(a[1], b[1])), TVM_MUL(a[2], b[2]));
}

// tensorBase.py:344-348 (uniform marching) / nerfplusplus.py:178-194,239-247 (sphere-bounded)
// `jitter` is [n] (uniform sampling, NULL = eval) or [n][S] (TVM_SAMPLING_NPP, required).
TVM_HD void ray_setup(const TvmModel& m, const float* ray6, const float* jitter, int ray, int S, RayMarch& r) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r.o[i] = ray6[i];
    r.d[i] = ray6[3 + i];
  }
  r.S = S;
  r.rnd = nullptr;
  r.step = 0.0f;
  if (m.sampling == TVM_SAMPLING_NPP) {
    // intersect_sphere(rays_o, rays_d, radii^2): d1 = -sum(d*o)/sum(d*d); p = o + d1 d; d2 = sqrt(R^2 - |p|^2) / |d|
    const float dd = dot3_seq(r.d, r.d);
    const float d1 = TVM_DIV(-dot3_seq(r.d, r.o), dd);
    float p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = TVM_ADD(r.o[i], TVM_MUL(d1, r.d[i]));
    const float cosd = TVM_DIV(1.0f, sqrtf(dd));
    const float d2 = TVM_MUL(sqrtf(TVM_SUB(TVM_MUL(m.radii, m.radii), dot3_seq(p, p))), cosd);
    const float far = TVM_ADD(d1, d2);
    r.t_min = m.near_;
    r.t_far = far;
    r.step = TVM_DIV(TVM_SUB(far, m.near_), (float)(S - 1));
    r.jit = 0.0f;
    r.rnd = jitter + (size_t)ray * S;
    return;
  }
  float t = -INFINITY, tf = INFINITY;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float vec = (r.d[i] == 0.0f) ? 1e-6f : r.d[i];
    float ra = TVM_DIV(TVM_SUB(m.aabb[3 + i], r.o[i]), vec);
    float rb = TVM_DIV(TVM_SUB(m.aabb[i], r.o[i]), vec);
    float mn = fminf(ra, rb);
    t = fmaxf(t, mn);
    tf = fminf(tf, fmaxf(ra, rb));
  }
  t = fminf(fmaxf(t, m.near_), m.far_);
  r.t_min = t;
  r.t_far = tf;
  r.jit = jitter ? jitter[ray] : 0.0f;
}

// z_k = t_min + stepSize * (k + jitter)                        (tensorBase.py:350-355)
// NeRF++: fg_i = near + i*step; stratum [lower_i, upper_i] from the mid points; z_i = lower + (upper-lower) t_i
//                                                              (nerfplusplus.py:196-205, 248-251)
TVM_HD float sample_z(const TvmModel& m, const RayMarch& r, int k) {
  if (m.sampling == TVM_SAMPLING_NPP) {
    const int kc = min(max(k, 0), r.S - 1);      // callers probe k = S for the last dist, which is unused
    const float f0 = TVM_ADD(r.t_min, TVM_MUL((float)kc, r.step));
    float lower = f0, upper = f0;
    if (kc > 0) lower = TVM_MUL(0.5f, TVM_ADD(f0, TVM_ADD(r.t_min, TVM_MUL((float)(kc - 1), r.step))));
    if (kc < r.S - 1) upper = TVM_MUL(0.5f, TVM_ADD(TVM_ADD(r.t_min, TVM_MUL((float)(kc + 1), r.step)), f0));
    return TVM_ADD(lower, TVM_MUL(TVM_SUB(upper, lower), r.rnd[kc]));
  }
  float rng = TVM_ADD((float)k, r.jit);
  return TVM_ADD(r.t_min, TVM_MUL(m.step_size, rng));
}
// the uniform branch alone, for kernels specialised on the sampling mode
TVM_HD float sample_z_uniform(const TvmModel& m, const RayMarch& r, int k) {
  return TVM_ADD(r.t_min, TVM_MUL(m.step_size, TVM_ADD((float)k, r.jit)));
}

// pts = o + d * z ; returns true when the point is inside the bbox (strict > on both faces)
TVM_HD bool sample_point(const TvmModel& m, const RayMarch& r, float z, float p[3]) {
  bool out = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    p[i] = TVM_ADD(r.o[i], TVM_MUL(r.d[i], z));
    out = out || (m.aabb[i] > p[i]) || (p[i] > m.aabb[3 + i]);
  }
  return !out;
}

// grid_sample index arithmetic (A1, align_corners=True): u = ((c + 1) / 2) * (size - 1)
TVM_HD float unnormalize(float c, int size) {
  return TVM_MUL(TVM_MUL(TVM_ADD(c, 1.0f), 0.5f), (float)(size - 1));
}

// One interpolation axis: lower/upper tap index (clamped into range) and their weights;
// a tap that falls outside [0,size-1] gets weight 0 (zeros padding).
struct Axis {
  int i0, i1;
  float w0, w1;
};
TVM_HD Axis axis_taps(float u, int size) {
  Axis a;
  float f = floorf(u);
  a.w0 = TVM_SUB(TVM_ADD(f, 1.0f), u);
  a.w1 = TVM_SUB(u, f);
  int i = (int)f;
  if (i < 0 || i > size - 1) a.w0 = 0.0f;
  if (i + 1 < 0 || i + 1 > size - 1) a.w1 = 0.0f;
  a.i0 = min(max(i, 0), size - 1);
  a.i1 = min(max(i + 1, 0), size - 1);
  return a;
}

// The same two taps re-expressed on the ADJACENT pair (b, b+1), b = clamp(floor(u), 0, size-2): p0 weighs texel b, p1
// texel b+1.  In range (0 <= floor(u) <= size-2) this is axis_taps verbatim; at floor(u) = size-1 (a sample exactly on the
// far face) the in-range tap moves to p1 and the zero-weight clamped tap to p0, at floor(u) = -1 the other way round, so
// p0 T[b] + p1 T[b+1] has the same non-zero products in the same order (the zero-weight terms add exact zeros).
// A gather then needs ONE address per row: the neighbour sits at a constant +C offset.  Requires size >= 2.
struct AxisPair {
  int b;
  float p0, p1;
};
TVM_HD AxisPair axis_pair(float u, int size) {
  AxisPair a;
  const float f = floorf(u);
  const float w0 = TVM_SUB(TVM_ADD(f, 1.0f), u), w1 = TVM_SUB(u, f);
  const int i = (int)f;
  a.b = min(max(i, 0), size - 2);
  a.p0 = (i == a.b) ? w0 : ((i + 1 == a.b) ? w1 : 0.0f);
  a.p1 = (i == a.b) ? w1 : ((i == a.b + 1) ? w0 : 0.0f);
  return a;
}

// AlphaGridMask.sample_alpha(...) > 0                          (tensorBase.py:50-59, 491-493)
// With a non-negative volume the trilinear value is > 0 iff some in-range corner with a non-zero
// product weight has its bit set (the product cannot underflow: each factor is 0 or >= 2^-25*(size-1)).
TVM_HD bool alpha_mask_test(const TvmModel& m, const uint32_t* __restrict__ bits, const float p[3]) {
  Axis ax[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float c = TVM_SUB(TVM_MUL(TVM_SUB(p[i], m.alpha_aabb_min[i]), m.alpha_inv_size[i]), 1.0f);
    ax[i] = axis_taps(unnormalize(c, m.alpha_grid[i]), m.alpha_grid[i]);
  }
  const int W = m.alpha_grid[0], H = m.alpha_grid[1];
  if (m.alpha_dilated != nullptr && ax[0].w0 > 0.0f && ax[0].w1 > 0.0f && ax[1].w0 > 0.0f && ax[1].w1 > 0.0f &&
      ax[2].w0 > 0.0f && ax[2].w1 > 0.0f) {
    // all 8 taps in range with non-zero weights (a zero weight only arises on a lattice plane or outside the volume):
    // the decision is the OR of the 8 corner bits, precomputed by tvm_pack_alpha_dilated
    const uint32_t idx = ((uint32_t)ax[2].i0 * (uint32_t)H + (uint32_t)ax[1].i0) * (uint32_t)W + (uint32_t)ax[0].i0;
    return (m.alpha_dilated[idx >> 5] >> (idx & 31u)) & 1u;
  }
  bool hit = false;
#pragma unroll
  for (int cz = 0; cz < 2; ++cz) {
    float wz = cz ? ax[2].w1 : ax[2].w0;
    int z = cz ? ax[2].i1 : ax[2].i0;
#pragma unroll
    for (int cy = 0; cy < 2; ++cy) {
      float wy = cy ? ax[1].w1 : ax[1].w0;
      int y = cy ? ax[1].i1 : ax[1].i0;
#pragma unroll
      for (int cx = 0; cx < 2; ++cx) {
        float wx = cx ? ax[0].w1 : ax[0].w0;
        int x = cx ? ax[0].i1 : ax[0].i0;
        bool w_nz = (wx > 0.0f) && (wy > 0.0f) && (wz > 0.0f);
        uint32_t idx = ((uint32_t)z * (uint32_t)H + (uint32_t)y) * (uint32_t)W + (uint32_t)x;
        uint32_t word = bits[idx >> 5];
        hit = hit || (w_nz && ((word >> (idx & 31u)) & 1u));
      }
    }
  }
  return hit;
}

// Conservative empty-space test for a run of consecutive samples whose end points are p0 and p1:
// the mask-space coordinate of every sample in between lies between the end points' (each step of the
// coordinate arithmetic is monotone in k), so the voxels any of them can touch are [min floor, max floor + 1]
// per axis.  Returns false only if every 8^3 brick overlapping that box is empty => no sample of the
// run can pass alpha_mask_test.  Used to skip whole 32-sample blocks; never changes a mask decision.
TVM_HD bool bricks_maybe(const TvmModel& m, const uint32_t* __restrict__ bricks, const float p0[3], const float p1[3]) {
  int lo[3], hi[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float c0 = TVM_SUB(TVM_MUL(TVM_SUB(p0[i], m.alpha_aabb_min[i]), m.alpha_inv_size[i]), 1.0f);
    const float c1 = TVM_SUB(TVM_MUL(TVM_SUB(p1[i], m.alpha_aabb_min[i]), m.alpha_inv_size[i]), 1.0f);
    const float f0 = floorf(unnormalize(c0, m.alpha_grid[i])), f1 = floorf(unnormalize(c1, m.alpha_grid[i]));
    const float a = fminf(f0, f1), b = fmaxf(f0, f1) + 1.0f;
    const float va = fmaxf(a, 0.0f), vb = fminf(b, (float)(m.alpha_grid[i] - 1));
    if (!(va <= vb)) return false;      // the run touches no voxel on this axis (also catches NaN)
    lo[i] = (int)va >> 3;
    hi[i] = (int)vb >> 3;
  }
  const int BW = (m.alpha_grid[0] + 7) >> 3, BH = (m.alpha_grid[1] + 7) >> 3;
  if (m.alpha_bricks3 != nullptr && hi[0] - lo[0] <= 2 && hi[1] - lo[1] <= 2 && hi[2] - lo[2] <= 2) {
    // the box spans at most 3 bricks per axis: it lies inside the 3x3x3 neighbourhood of its middle brick, whose occupancy is
    // one 27-bit word (tvm_pack_alpha_bricks3); the box is a product of three per-axis masks inside that word
    const int mid[3] = {(lo[0] + hi[0]) >> 1, (lo[1] + hi[1]) >> 1, (lo[2] + hi[2]) >> 1};
    const uint32_t word = m.alpha_bricks3[((uint32_t)mid[2] * BH + mid[1]) * BW + mid[0]];
    if (word == 0u) return false;
    uint32_t ax[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) ax[i] = ((1u << (hi[i] - lo[i] + 1)) - 1u) << (lo[i] - mid[i] + 1);      // bits d + 1, d in [-1, 1]
    const uint32_t sy = (ax[1] & 1u) | ((ax[1] & 2u) << 2) | ((ax[1] & 4u) << 4);       // y mask spread to bits 0, 3, 6
    const uint32_t sz = (ax[2] & 1u) | ((ax[2] & 2u) << 8) | ((ax[2] & 4u) << 16);      // z mask spread to bits 0, 9, 18
    return (word & (ax[0] * sy * sz)) != 0u;                                            // products without carries
  }
  for (int z = lo[2]; z <= hi[2]; ++z)
    for (int y = lo[1]; y <= hi[1]; ++y)
      for (int x = lo[0]; x <= hi[0]; ++x) {
        const uint32_t idx = ((uint32_t)z * BH + y) * BW + x;
        if ((bricks[idx >> 5] >> (idx & 31u)) & 1u) return true;
      }
  return false;
}

// Which 32-sample blocks of a ray can contain a sample that passes the bbox test and the alpha mask?
// Block b is dropped when both of its end samples lie beyond the same bbox face (monotonicity => so do
// all samples in between) or when bricks_maybe() is false.  Lane l decides block b0 + l.
// `uniform` (a compile-time constant at the call site) promises m.sampling != TVM_SAMPLING_NPP.
TVM_HD bool block_maybe(const TvmModel& m, const RayMarch& r, int b, int S, bool uniform = false) {
  const int k0 = b * 32, k1 = min(b * 32 + 31, S - 1);
  float p0[3], p1[3];
  sample_point(m, r, uniform ? sample_z_uniform(m, r, k0) : sample_z(m, r, k0), p0);
  sample_point(m, r, uniform ? sample_z_uniform(m, r, k1) : sample_z(m, r, k1), p1);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if ((p0[i] < m.aabb[i] && p1[i] < m.aabb[i]) || (p0[i] > m.aabb[3 + i] && p1[i] > m.aabb[3 + i])) return false;
  }
  if (m.alpha_bits != nullptr && m.alpha_bricks != nullptr) return bricks_maybe(m, m.alpha_bricks, p0, p1);
  return true;
}

// normalize_coord + unnormalize for the model grids             (tensorBase.py:223-224)
TVM_HD void grid_coords(const TvmModel& m, const float p[3], float u[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float c = TVM_SUB(TVM_MUL(TVM_SUB(p[i], m.aabb[i]), m.inv_aabb_size[i]), 1.0f);
    u[i] = unnormalize(c, m.grid[i]);
  }
}

// feature2density                                               (tensorBase.py:444-448; A3)
TVM_HD float feature2density(const TvmModel& m, float f) {
  if (m.act == TVM_ACT_RELU) return fmaxf(f, 0.0f);
  float x = f + m.density_shift;
  return logf(1.0f + expf(fminf(x, 20.0f))) + fmaxf(x - 20.0f, 0.0f);
}
// d softplus / dx of the formula above
TVM_HD float feature2density_grad(const TvmModel& m, float f) {
  if (m.act == TVM_ACT_RELU) return f > 0.0f ? 1.0f : 0.0f;
  float x = f + m.density_shift;
  if (x > 20.0f) return 1.0f;
  float e = expf(x);
  return e / (1.0f + e);
}

// 16-byte read-only load (texel segments are 16-byte aligned: channel counts are multiples of 4)
TVM_HD float4 ldg4(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(reinterpret_cast<const float4*>(p));
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}

// Bilinear plane x linear line product for 4 consecutive channels starting at c.
struct VmTaps {
  int o00, o01, o10, o11;   // plane texel offsets (in texels)
  float nw, ne, sw, se;     // (x0,y0) (x1,y0) (x0,y1) (x1,y1) weights, products formed first (A1)
  int l0, l1;               // line rows
  float lw0, lw1;
};
TVM_HD VmTaps vm_taps(const TvmModel& m, const Axis ax[3], int k) {
  const Axis& aw = ax[mat0(k)];
  const Axis& ah = ax[mat1(k)];
  const Axis& al = ax[vecax(k)];
  const int W = m.grid[mat0(k)];
  VmTaps t;
  t.o00 = ah.i0 * W + aw.i0;
  t.o01 = ah.i0 * W + aw.i1;
  t.o10 = ah.i1 * W + aw.i0;
  t.o11 = ah.i1 * W + aw.i1;
  t.nw = TVM_MUL(aw.w0, ah.w0);
  t.ne = TVM_MUL(aw.w1, ah.w0);
  t.sw = TVM_MUL(aw.w0, ah.w1);
  t.se = TVM_MUL(aw.w1, ah.w1);
  t.l0 = al.i0;
  t.l1 = al.i1;
  t.lw0 = al.w0;
  t.lw1 = al.w1;
  return t;
}
TVM_HD void vm_sample4(const float* __restrict__ plane, const float* __restrict__ line,
                                           const VmTaps& t, int C, int c, float4& pv, float4& lv) {
  float4 a = ldg4(plane + (size_t)t.o00 * C + c);
  float4 b = ldg4(plane + (size_t)t.o01 * C + c);
  float4 cc = ldg4(plane + (size_t)t.o10 * C + c);
  float4 d = ldg4(plane + (size_t)t.o11 * C + c);
  float4 l0 = ldg4(line + (size_t)t.l0 * C + c);
  float4 l1 = ldg4(line + (size_t)t.l1 * C + c);
  pv.x = a.x * t.nw + b.x * t.ne + cc.x * t.sw + d.x * t.se;
  pv.y = a.y * t.nw + b.y * t.ne + cc.y * t.sw + d.y * t.se;
  pv.z = a.z * t.nw + b.z * t.ne + cc.z * t.sw + d.z * t.se;
  pv.w = a.w * t.nw + b.w * t.ne + cc.w * t.sw + d.w * t.se;
  lv.x = l0.x * t.lw0 + l1.x * t.lw1;
  lv.y = l0.y * t.lw0 + l1.y * t.lw1;
  lv.z = l0.z * t.lw0 + l1.z * t.lw1;
  lv.w = l0.w * t.lw0 + l1.w * t.lw1;
}

// Pair form of vm_taps + vm_sample4 for 4 consecutive channels starting at c (same expressions, hence the same
// FMA contraction and the same bits): two row addresses per plane and one per line, neighbours at +C.
struct VmPair {
  uint32_t row0, row1, lrow;   // element offsets of (h, w), (h+1, w) in the plane and of row l in the line
  float nw, ne, sw, se, lw0, lw1;
};
TVM_HD VmPair vm_pair(const TvmModel& m, const AxisPair ax[3], int k, int C) {
  const AxisPair& aw = ax[mat0(k)];
  const AxisPair& ah = ax[mat1(k)];
  const AxisPair& al = ax[vecax(k)];
  const uint32_t W = (uint32_t)m.grid[mat0(k)];
  VmPair t;
  t.row0 = ((uint32_t)ah.b * W + (uint32_t)aw.b) * (uint32_t)C;
  t.row1 = t.row0 + W * (uint32_t)C;
  t.lrow = (uint32_t)al.b * (uint32_t)C;
  t.nw = TVM_MUL(aw.p0, ah.p0);
  t.ne = TVM_MUL(aw.p1, ah.p0);
  t.sw = TVM_MUL(aw.p0, ah.p1);
  t.se = TVM_MUL(aw.p1, ah.p1);
  t.lw0 = al.p0;
  t.lw1 = al.p1;
  return t;
}
TVM_HD void vm_pair_sample4(const float* __restrict__ plane, const float* __restrict__ line, const VmPair& t, int C, int c,
                            float4& pv, float4& lv) {
  const float* r0 = plane + t.row0 + c;
  const float* r1 = plane + t.row1 + c;
  const float* lr = line + t.lrow + c;
  float4 a = ldg4(r0);
  float4 b = ldg4(r0 + C);
  float4 cc = ldg4(r1);
  float4 d = ldg4(r1 + C);
  float4 l0 = ldg4(lr);
  float4 l1 = ldg4(lr + C);
  pv.x = a.x * t.nw + b.x * t.ne + cc.x * t.sw + d.x * t.se;
  pv.y = a.y * t.nw + b.y * t.ne + cc.y * t.sw + d.y * t.se;
  pv.z = a.z * t.nw + b.z * t.ne + cc.z * t.sw + d.z * t.se;
  pv.w = a.w * t.nw + b.w * t.ne + cc.w * t.sw + d.w * t.se;
  lv.x = l0.x * t.lw0 + l1.x * t.lw1;
  lv.y = l0.y * t.lw0 + l1.y * t.lw1;
  lv.z = l0.z * t.lw0 + l1.z * t.lw1;
  lv.w = l0.w * t.lw0 + l1.w * t.lw1;
}

}  // namespace tvm
