// Shared declarations of libtvmrender's translation units (workspace carve-up, launch params).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>
#include "tvm_math.cuh"

namespace tvm {

void set_error(const char* fmt, ...);

#define TVM_CHECK_CUDA(expr)                                                         \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      tvm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                     \
    }                                                                                \
  } while (0)

#define TVM_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      tvm::set_error(__VA_ARGS__);    \
      return -1;                      \
    }                                 \
  } while (0)

constexpr int kMaxAppDim = 32;     // basis outputs are padded to 32 columns
constexpr int kFeatureC = 128;     // MLP width the kernels are specialised for
constexpr float kErtEps = 1e-7f;   // early ray termination: remaining weight mass < 1e-7

// Scratch layout for n rays x S samples (all offsets 256-byte aligned).  Entries are the samples
// with weight > thres, appended block-of-32-samples at a time; (blk_mask, blk_base) map a
// (ray, sample-block) back to its entries so that compositing and the backward pass stay in
// ray/sample order without a sort.
struct Workspace {
  uint32_t* n_entries;   // [1] (+ padding)
  uint32_t* blk_mask;    // [n][NB]   app_mask bits of each 32-sample block
  uint32_t* blk_base;    // [n][NB]   first entry index of the block
  uint2* ent;            // [cap]     (ray, sample index)
  float* ent_w;          // [cap]     weight
  float4* ent_u;         // [cap]     un-normalised grid coordinates of the sample (x, y, z) and its weight: the appearance
                         //           gather starts from one 16-byte load instead of re-deriving the ray march
  float* ent_rgb;        // [cap][3]  per-sample colour written by the appearance stage
  float* ent_pen;        // [cap]     TVM_VARIANT_REF: relu(-d.n)^2 of the sample (REFTensoRF.py:236-237)
  float* pen_sum;        // [n]       TVM_VARIANT_REF: sum_k w_k * pen_k of the ray
  float* acc;            // [n]       acc_map
  float* rgb_sum;        // [n][3]    sum_s w * rgb (before white background / clamp)
  float* bwd_scratch;    // [n][4]    per-ray scratch of the backward pass
  float* bg_lambda;      // [n]       NeRF++: prod(1 - alpha + 1e-6) of the foreground, 0 when <= 0.1
  uint32_t* bg_list;     // [n]       NeRF++: rays whose background is evaluated (bg_lambda > 0.1); count at n_entries[1]
  float* bg_rgb;         // [n][3]    NeRF++: composited background colour of the evaluated rays (before the bg_lambda factor)
  float* bwd_lam;        // [n]       NeRF++ backward: bg_lambda * dL/d bg_lambda of the ray
  uint32_t cap;          // entries the arrays hold: n * S, or fewer in a bounded workspace (then *n_entries > cap = overflow)
  int NB;
  size_t bytes;
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// `max_entries` bounds the entry list below the worst case n * S (tvm_forward with a workspace smaller than
// tvm_workspace_bytes): every per-entry array is cut to that capacity, the per-ray tables keep their size.
inline Workspace carve_workspace(void* base, int n, int S, size_t max_entries = ~size_t(0)) {
  Workspace w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align256(bytes);
    return r;
  };
  w.NB = (S + 31) / 32;
  w.cap = (uint32_t)std::min((size_t)n * (size_t)S, max_entries);
  w.n_entries = (uint32_t*)take(256);
  w.blk_mask = (uint32_t*)take((size_t)n * w.NB * 4);
  w.blk_base = (uint32_t*)take((size_t)n * w.NB * 4);
  w.ent = (uint2*)take((size_t)w.cap * 8);
  w.ent_w = (float*)take((size_t)w.cap * 4);
  w.ent_u = (float4*)take((size_t)w.cap * 16);
  w.ent_rgb = (float*)take((size_t)w.cap * 12);
  w.ent_pen = (float*)take((size_t)w.cap * 4);
  w.pen_sum = (float*)take((size_t)n * 4);
  w.acc = (float*)take((size_t)n * 4);
  w.rgb_sum = (float*)take((size_t)n * 12);
  w.bwd_scratch = (float*)take((size_t)n * 16);
  w.bg_lambda = (float*)take((size_t)n * 4);
  w.bg_list = (uint32_t*)take((size_t)n * 4);
  w.bg_rgb = (float*)take((size_t)n * 12);
  w.bwd_lam = (float*)take((size_t)n * 4);
  w.bytes = off;
  return w;
}

// Largest entry capacity (<= n * S) whose carve-up fits ws_bytes; 0 when not even the per-ray tables fit.
inline uint32_t workspace_capacity(int n, int S, size_t ws_bytes) {
  const size_t worst = (size_t)n * (size_t)S;
  if (carve_workspace(nullptr, n, S).bytes <= ws_bytes) return (uint32_t)worst;
  const size_t fixed = carve_workspace(nullptr, n, S, 0).bytes;
  if (ws_bytes <= fixed) return 0;
  constexpr size_t kEntryBytes = 8 + 4 + 16 + 12 + 4;       // ent, ent_w, ent_u, ent_rgb, ent_pen
  size_t cap = std::min(worst, (ws_bytes - fixed) / kEntryBytes);
  while (cap > 0 && carve_workspace(nullptr, n, S, cap).bytes > ws_bytes) --cap;     // the 256-byte padding of the five arrays: < 30 steps
  return (uint32_t)cap;
}

struct FwdParams {
  TvmModel m;
  const float* rays;
  const float* jitter;
  int n, S, NB;
  uint32_t flags;
  Workspace ws;
  float* rgb_map;
  float* depth_map;
  TvmAux aux;
  unsigned long long* counters;
  int in_mlp_c;   // 2*view_pe*3 + 2*fea_pe*app_dim + 3 + app_dim (+1 for TVM_VARIANT_REF)
  int st;         // smem row stride (floats) of the fp32 appearance tiles
  // NeRF++ background (tvm_forward_npp)
  const float* bg_rand;   // [n][512]
  TvmBgNet bg;
};

inline int in_mlp_c(const TvmModel& m) {
  return 2 * m.view_pe * 3 + 2 * m.fea_pe * m.app_dim + 3 + m.app_dim + (m.variant == TVM_VARIANT_REF ? 1 : 0);
}
// MLP input column layout (tensorBase.py:76-83 / REFTensoRF.py:19-25): [(-dot)] feat dir PE(feat) PE(dir)
__host__ __device__ inline int col_feat(const TvmModel& m) { return m.variant == TVM_VARIANT_REF ? 1 : 0; }
__host__ __device__ inline int col_dir(const TvmModel& m) { return col_feat(m) + m.app_dim; }
__host__ __device__ inline int head_ld(const TvmModel& m) { return m.variant == TVM_VARIANT_REF ? TVM_REF_HEAD_LD : 32; }

int validate_model(const TvmModel& m);

// measurement hooks (tvm_profile_enable): RAII bracket of one kernel launch with cudaEvents
bool profile_on();
void profile_begin(int stage, cudaStream_t s);
void profile_end(cudaStream_t s);
struct ProfileScope {
  cudaStream_t s;
  bool on;
  ProfileScope(int stage, cudaStream_t s_) : s(s_), on(profile_on()) { if (on) profile_begin(stage, s); }
  ~ProfileScope() { if (on) profile_end(s); }
};

// TVM_EVAL_ONLY: per-ray sums of w * rgb in 32-bit FIXED POINT, three words per ray, in the space of the (unused) per-block
// tables.  Units of 2^-31: the colours of TVM_VARIANT_VM are sigmoids in (0, 1) and sum(w) <= 1, so a ray's sum stays below
// 2^31; a term is rounded to 4.7e-10 (fp32 itself resolves 6e-8 at 1), a ray of 1036 terms to < 2.5e-7 in the worst case.
// Integer addition is associative: the sum does not depend on the order in which the entries arrive.
// (64-bit sums in 2^-36 units measured 0.03 ms per frame slower: F2I.S64 is emulated, and the 64-bit REDs cost more.)
constexpr float kFixScale = 2147483648.0f;             // 2^31
constexpr float kFixInv = 1.0f / 2147483648.0f;
__host__ __device__ inline uint32_t* fix_sums(const Workspace& w) { return w.blk_mask; }
inline bool fix_sums_fit(int n, int NB) { return NB >= 3; }      // 12 bytes per ray inside n * NB * 4
#if defined(__CUDACC__)
__device__ __forceinline__ void fix_accumulate(uint32_t* sums, uint32_t ray, float w, float r, float g, float b) {
  uint32_t* a = sums + 3 * (size_t)ray;
  atomicAdd(a + 0, __float2uint_rn(w * r * kFixScale));       // result unused: RED.E.ADD
  atomicAdd(a + 1, __float2uint_rn(w * g * kFixScale));
  atomicAdd(a + 2, __float2uint_rn(w * b * kFixScale));
}
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Entry (ray, k) -> un-normalised grid coordinates, with exactly the arithmetic of the march kernel.
__device__ __forceinline__ void entry_coords(const TvmModel& m, const float* rays, const float* jitter,
                                             uint32_t ray, uint32_t k, int S, float u[3], float dir[3]) {
  RayMarch r;
  ray_setup(m, rays + 6 * (size_t)ray, jitter, (int)ray, S, r);
  float z = sample_z(m, r, (int)k);
  float p[3];
  sample_point(m, r, z, p);
  grid_coords(m, p, u);
  dir[0] = r.d[0];
  dir[1] = r.d[1];
  dir[2] = r.d[2];
}

}  // namespace tvm
