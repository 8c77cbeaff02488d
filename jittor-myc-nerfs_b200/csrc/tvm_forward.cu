// Forward pass of the TensoRF-VM ray renderer (TensorBase.execute, tensorf-myc/models/tensorBase.py:476-536).
//
//   k_march      one warp per ray, lanes = 32 consecutive samples: sample_ray (:340-360), bbox mask,
//                AlphaGridMask test on the bit-packed volume (:39-59), density gather over the three
//                plane x line factors (tensoRF.py:209-225), softplus (:444-448), raw2alpha with a warp
//                prefix product (:17-24), weight threshold (:513), acc/depth reductions (:520-531), early
//                ray termination, and appends the weighted samples to the appearance entry list.
//   k_app_simt   appearance head in fp32 for a tile of 64 entries: 48-channel gathers
//                (tensoRF.py:228-244), basis_mat, positional encoding and MLPRender_Fea (:62-86, 9-15).
//   k_composite  one warp per ray, lane = sample of a block: rgb_map = clamp(sum w*rgb + (1-acc)) (:521-528).
//
// The tcgen05 tensor-core appearance head lives in tvm_mlp_tc.cu and replaces k_app_simt when
// TVM_MLP_BF16 / TVM_MLP_FP16 is requested.
#include <atomic>
#include "tvm_app_simt.cuh"

namespace tvm {

#ifndef TVM_MARCH_WARPS
#define TVM_MARCH_WARPS 4
#endif
constexpr int kMarchWarps = TVM_MARCH_WARPS;      // 128-thread CTAs: a CTA lives as long as its slowest ray, smaller CTAs pack better (8 -> 4 warps: -2.4 %)
#ifndef TVM_MARCH_MIN_CTAS
#define TVM_MARCH_MIN_CTAS 8      // 64 registers, 32 warps per SM (measured: 24 -> 32 warps = -8 % march time)
#endif

// ------------------------------------------------------------------------------------------------
// k_march
// ------------------------------------------------------------------------------------------------
// AUX: parity instantiation (every mask bit and per-sample output written, no early termination, no block skipping).
// NPP: NerfPlusPlus sampling (sphere-bounded, stratified) -- the uniform instantiations carry none of that code.
// CD:  compile-time density channel count (16 = configs/*.txt; 0 = any multiple of 4, read from the model).
template <bool AUX, bool NPP, int CD>
__global__ void __launch_bounds__(kMarchWarps * 32, TVM_MARCH_MIN_CTAS) k_march(const FwdParams P) {
  // (measured and dropped, profiles/r02_notes.txt: computing the three axis_pair of a sample once in its own lane and
  //  handing them to its four gather lanes through 48 B of shared memory per sample -- 10 % fewer instructions, but 3.6 % SLOWER)
  __shared__ float s_u[kMarchWarps][32][3];
  __shared__ float s_f[kMarchWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kMarchWarps + warp;
  if (ray >= P.n) return;
  const TvmModel& m = P.m;
  const uint32_t lt_mask = (1u << lane) - 1u;

  RayMarch r;
  ray_setup(m, P.rays + 6 * (size_t)ray, P.jitter, ray, P.S, r);
  auto z_of = [&](int k) { return NPP ? sample_z(m, r, k) : sample_z_uniform(m, r, k); };

  const int C = CD ? CD : m.n_density;
  const int S = P.S;
  const bool ert = !AUX && !(P.flags & TVM_NO_ERT);
  float T = 1.0f, acc = 0.0f, dep = 0.0f;
  // NeRF++: bg_lambda = prod_k (1 - alpha_k + 1e-6) over ALL S samples (nerfplusplus.py:277-278)
  float Tbg = 1.0f;
  int n_proc = 0;
  bool seen = false;
  uint32_t c_in = 0, c_v = 0, c_a = 0;

  // Blocks are visited in windows of 32: a coarse pass (one lane per 32-sample block) drops the blocks of the window that
  // cannot hold a valid sample.  Uniform marching: the ray leaves the box at t_far (slab test), so windows that start
  // more than two samples beyond it are never looked at (rounding in o + d z is orders of magnitude below one step).
  int nb_hi = P.NB;
  if (!AUX && !NPP) {
    // ray_setup's slab test also yields the exit distance
    const float kf = __fdividef(r.t_far - r.t_min, m.step_size) + 3.0f;
    if (kf < (float)S) nb_hi = min(P.NB, max(0, (int)kf) / 32 + 1);     // NaN / inf keep NB
  }
  uint32_t visit = 0;
  int win = -32;
  int b = -1;
  while (true) {
    if (!AUX) {
      while (!visit) {
        win += 32;
        if (win >= nb_hi) break;
        const int bb = win + lane;
        visit = __ballot_sync(0xffffffffu, bb < nb_hi && block_maybe(m, r, bb, S, !NPP));
      }
      if (!visit) break;
      b = win + __ffs(visit) - 1;
      visit &= visit - 1;
    } else if (++b >= P.NB) {
      break;
    }
    const int k = b * 32 + lane;
    const float z = z_of(k);
    float p[3];
    bool inside = sample_point(m, r, z, p) && (k < S);
    const uint32_t in_bits = __ballot_sync(0xffffffffu, inside);
    if (AUX && lane == 0 && P.aux.bbox_bits) P.aux.bbox_bits[(size_t)ray * P.NB + b] = in_bits;
    if (in_bits == 0) {
      // per-axis monotonicity of o + d*z in k makes the in-box samples one contiguous run:
      // an empty block after a non-empty one means the ray has left the box for good.
      if (seen) break;
      continue;
    }
    seen = true;
    c_in += __popc(in_bits);

    bool valid = inside;
    if (m.alpha_bits != nullptr && inside) valid = alpha_mask_test(m, m.alpha_bits, p);
    const uint32_t v_bits = __ballot_sync(0xffffffffu, valid);
    if (AUX && lane == 0 && P.aux.valid_bits) P.aux.valid_bits[(size_t)ray * P.NB + b] = v_bits;

    if (!AUX && !NPP && v_bits == 0) {
      // no density anywhere in the block: alpha = 0, T / acc / depth unchanged, nothing to append
      if (!(in_bits >> 31)) break;
      continue;
    }
    float sigma = 0.0f;
    if (v_bits != 0) {
      const int nv = __popc(v_bits);
      c_v += nv;
      const int rank = __popc(v_bits & lt_mask);
      if (valid) {
        float u[3];
        grid_coords(m, p, u);
        s_u[warp][rank][0] = u[0];
        s_u[warp][rank][1] = u[1];
        s_u[warp][rank][2] = u[2];
      }
      __syncwarp();
      // 4 lanes per sample, one float4 of channels each: a tap is one 64-byte segment; the two rows of a plane
      // footprint and the line pair need one address each (axis_pair: the neighbour texel sits at +C)
      const int q = lane & 3;
      for (int g = 0; g < nv; g += 8) {
        const int j = g + (lane >> 2);
        float part = 0.0f;
        if (j < nv) {
          AxisPair ax[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) ax[i] = axis_pair(s_u[warp][j][i], m.grid[i]);
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const VmPair t = vm_pair(m, ax, kk, C);
            if (CD == 16) {
              float4 pv, lv;
              vm_pair_sample4(m.density_plane[kk], m.density_line[kk], t, 16, q * 4, pv, lv);
              part += pv.x * lv.x + pv.y * lv.y + pv.z * lv.z + pv.w * lv.w;
            } else {
              for (int c = q * 4; c < C; c += 16) {
                float4 pv, lv;
                vm_pair_sample4(m.density_plane[kk], m.density_line[kk], t, C, c, pv, lv);
                part += pv.x * lv.x + pv.y * lv.y + pv.z * lv.z + pv.w * lv.w;
              }
            }
          }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (j < nv && q == 0) s_f[warp][j] = part;
      }
      __syncwarp();
      if (valid) sigma = feature2density(m, s_f[warp][rank]);
      __syncwarp();
    }

    // raw2alpha: dists = z[k+1]-z[k] (last sample 0), scaled by distance_scale (:488, :511)
    const float z1 = z_of(k + 1);
    const float dist = (k < S - 1) ? TVM_MUL(TVM_SUB(z1, z), m.distance_scale) : 0.0f;
    const float alpha = TVM_SUB(1.0f, expf(TVM_MUL(-sigma, dist)));
    const float v = TVM_ADD(TVM_SUB(1.0f, alpha), 1e-10f);
    float pref = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, pref, o);
      if (lane >= o) pref *= t;
    }
    float excl = __shfl_up_sync(0xffffffffu, pref, 1);
    if (lane == 0) excl = 1.0f;
    const float w = alpha * (T * excl);
    T = T * __shfl_sync(0xffffffffu, pref, 31);
    if (NPP) {
      float v6 = (k < S) ? TVM_ADD(TVM_SUB(1.0f, alpha), 1e-6f) : 1.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v6 *= __shfl_xor_sync(0xffffffffu, v6, o);
      Tbg *= v6;
      n_proc += min(32, S - b * 32);
    }
    acc += w;
    dep += w * z;

    const bool app = w > m.weight_thres;
    const uint32_t a_bits = __ballot_sync(0xffffffffu, app);
    if (a_bits != 0) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(P.ws.n_entries, (uint32_t)__popc(a_bits));
      base = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t e = base + __popc(a_bits & lt_mask);
      if (app && e < P.ws.cap) {      // bounded workspace: entries behind its capacity are counted, not stored (tvm_forward_entries)
        P.ws.ent[e] = make_uint2((uint32_t)ray, (uint32_t)k);
        P.ws.ent_w[e] = w;
        float u[3];
        grid_coords(m, p, u);
        P.ws.ent_u[e] = make_float4(u[0], u[1], u[2], w);
      }
      if (lane == 0 && (AUX || !(P.flags & TVM_EVAL_ONLY))) {      // TVM_EVAL_ONLY: nobody reads the per-block tables
        P.ws.blk_mask[(size_t)ray * P.NB + b] = a_bits;
        P.ws.blk_base[(size_t)ray * P.NB + b] = base;
      }
      c_a += __popc(a_bits);
    }
    if (AUX) {
      if (k < S) {
        if (P.aux.sigma) P.aux.sigma[(size_t)ray * S + k] = sigma;
        if (P.aux.weight) P.aux.weight[(size_t)ray * S + k] = w;
      }
      if (lane == 0 && P.aux.app_bits) P.aux.app_bits[(size_t)ray * P.NB + b] = a_bits;
    }
    if (ert && T < kErtEps) break;
    if (!(in_bits >> 31)) break;  // last lane already outside: the run of in-box samples has ended
  }

  acc = warp_sum(acc);
  dep = warp_sum(dep);
  if (lane == 0) {
    P.ws.acc[ray] = acc;
    if (NPP) {
      // samples of blocks that were never visited have alpha = 0: each contributes fl(1 + 1e-6)
      float lam = Tbg * powf(TVM_ADD(1.0f, 1e-6f), (float)(S - n_proc));
      lam = lam > 0.1f ? lam : 0.0f;                                   // nerfplusplus.py:313
      P.ws.bg_lambda[ray] = lam;
      if (lam > 0.0f) P.ws.bg_list[atomicAdd(P.ws.n_entries + 1, 1u)] = (uint32_t)ray;
      if (P.aux.bg_lambda) P.aux.bg_lambda[ray] = lam;
    }
    // depth_map = sum(w*z) + (1-acc) * rays_chunk[..., -1]  -- column 5 (d_z), reference quirk (:531)
    P.depth_map[ray] = dep + (1.0f - acc) * r.d[2];
    if (AUX && P.aux.acc_map) P.aux.acc_map[ray] = acc;
    if (P.counters) {
      atomicAdd(&P.counters[TVM_CNT_M_IN], (unsigned long long)c_in);
      atomicAdd(&P.counters[TVM_CNT_M_V], (unsigned long long)c_v);
      atomicAdd(&P.counters[TVM_CNT_M_A], (unsigned long long)c_a);
      if (c_v) atomicAdd(&P.counters[TVM_CNT_RAYS], 1ull);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_app_simt: appearance head, fp32 FMA path (parity mode); building blocks in tvm_app_simt.cuh
// ------------------------------------------------------------------------------------------------
template <int NH>
__global__ void __launch_bounds__(kAppThreads) k_app_simt(const FwdParams P) {
  extern __shared__ __align__(16) float smem[];
  const int st = P.st;
  float* H = smem;                    // [64][st]: appearance vector, then layer-1 output
  float* X = smem + kAppTile * st;    // [64][st]: MLP input, then layer-2 output
  float* HD = smem + 2 * kAppTile * st;   // [64][8]: REFTensoRF head outputs -> {rgb_d, tint}
  const TvmModel& m = P.m;
  const uint32_t n_ent = min(*P.ws.n_entries, P.ws.cap);
  const uint32_t n_tiles = (n_ent + kAppTile - 1) / kAppTile;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t tile_base = tile * kAppTile;
    app_gather_tile(P, tile_base, n_ent, H, X, st);
    __syncthreads();
    app_basis_pe<NH>(P, H, X, HD, st, tile_base, n_ent);
    __syncthreads();
    app_dense<true>(m.w1_t, m.b1, X, P.in_mlp_c, H, st);
    __syncthreads();
    app_dense<true>(m.w2_t, m.b2, H, kFeatureC, X, st);
    __syncthreads();
    {
      const int row = threadIdx.x & (kAppTile - 1), part = threadIdx.x >> 6;
      const uint32_t e = tile_base + row;
      if (part < 3 && e < n_ent) {
        float c = 1.0f / (1.0f + expf(-app_out_logit(m, X, st, row, part)));
        // REFTensoRF.py:232: rgb = specular_tint * clamp(rgb_s, 0) + rgb_d
        if (NH == TVM_REF_HEAD_LD) c = HD[row * 8 + 3] * fmaxf(c, 0.0f) + HD[row * 8 + part];
        if (P.flags & TVM_EVAL_ONLY)       // composite as we go: w * rgb into the ray's fixed-point sum
          atomicAdd(fix_sums(P.ws) + 3 * (size_t)P.ws.ent[e].x + part, __float2uint_rn(P.ws.ent_w[e] * c * kFixScale));
        else
          P.ws.ent_rgb[(size_t)e * 3 + part] = c;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// k_composite
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_composite(const FwdParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * 8 + warp;
  if (ray >= P.n) return;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, sp = 0.0f;
  const bool ref = P.m.variant == TVM_VARIANT_REF;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // lanes read the block table 32 blocks at a time; each non-empty block is then summed by the whole warp (lane = sample):
  // one round trip per block instead of one per entry; the lane sums meet in warp_sum's fixed tree, so the result does not
  // depend on chunking or launch order
  const uint32_t* __restrict__ row_mask = P.ws.blk_mask + (size_t)ray * P.NB;
  uint32_t nxt_bits = lane < P.NB ? __ldcs(row_mask + lane) : 0u;          // the table is read once: evict-first
  for (int b0 = 0; b0 < P.NB; b0 += 32) {
    const int bl = b0 + lane;
    const uint32_t my_bits = nxt_bits;
    if (b0 + 32 < P.NB) nxt_bits = bl + 32 < P.NB ? __ldcs(row_mask + bl + 32) : 0u;   // next round under this one's work
    const uint32_t my_base = my_bits ? __ldcs(P.ws.blk_base + (size_t)ray * P.NB + bl) : 0u;
    uint32_t todo = __ballot_sync(0xffffffffu, my_bits != 0u);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t bits = __shfl_sync(0xffffffffu, my_bits, src);
      const uint32_t base = __shfl_sync(0xffffffffu, my_base, src);
      const uint32_t e = base + __popc(bits & lt_mask);
      if (((bits >> lane) & 1u) && e < P.ws.cap) {
        const float w = P.ws.ent_w[e];
        const float r = P.ws.ent_rgb[(size_t)e * 3 + 0];
        const float g = P.ws.ent_rgb[(size_t)e * 3 + 1];
        const float bl_ = P.ws.ent_rgb[(size_t)e * 3 + 2];
        s0 = fmaf(w, r, s0);
        s1 = fmaf(w, g, s1);
        s2 = fmaf(w, bl_, s2);
        if (ref) sp = fmaf(w, P.ws.ent_pen[e], sp);
        if (P.aux.rgb) {
          float* o = P.aux.rgb + ((size_t)ray * P.S + (size_t)(b0 + src) * 32 + lane) * 3;
          o[0] = r; o[1] = g; o[2] = bl_;
        }
      }
    }
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (ref) sp = warp_sum(sp);
  if (lane == 0) {
    if (ref) P.ws.pen_sum[ray] = sp;
    const float acc = P.ws.acc[ray];
    P.ws.rgb_sum[(size_t)ray * 3 + 0] = s0;
    P.ws.rgb_sum[(size_t)ray * 3 + 1] = s1;
    P.ws.rgb_sum[(size_t)ray * 3 + 2] = s2;
    const float bg = (P.flags & TVM_WHITE_BG) ? (1.0f - acc) : 0.0f;
    P.rgb_map[(size_t)ray * 3 + 0] = fminf(fmaxf(s0 + bg, 0.0f), 1.0f);
    P.rgb_map[(size_t)ray * 3 + 1] = fminf(fmaxf(s1 + bg, 0.0f), 1.0f);
    P.rgb_map[(size_t)ray * 3 + 2] = fminf(fmaxf(s2 + bg, 0.0f), 1.0f);
  }
}

// TVM_EVAL_ONLY: the appearance head has summed w * rgb per ray in fixed point; convert, add the background, clamp (:521-528)
__global__ void __launch_bounds__(256) k_finalize(const FwdParams P) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= P.n) return;
  const uint32_t* a = fix_sums(P.ws) + 3 * (size_t)ray;
  const float bg = (P.flags & TVM_WHITE_BG) ? (1.0f - P.ws.acc[ray]) : 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = (float)a[c] * kFixInv;
    P.rgb_map[(size_t)ray * 3 + c] = fminf(fmaxf(s + bg, 0.0f), 1.0f);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int launch_app_tc(const FwdParams& P, int num_sms, cudaStream_t stream);   // tvm_mlp_tc.cu

int fill_fwd_params(FwdParams& P, const TvmModel* m, const float* rays, int n, int S, const float* jitter,
                    uint32_t flags, void* ws, size_t ws_bytes, bool bounded_ok) {
  TVM_REQUIRE(m && rays && ws, "null argument");
  TVM_REQUIRE(n > 0 && S > 0, "n_rays and n_samples must be positive");
  if (int rc = validate_model(*m)) return rc;
  P.m = *m;
  P.rays = rays;
  P.jitter = jitter;
  P.n = n;
  P.S = S;
  P.flags = flags;
  TVM_REQUIRE((double)n * S < 4.0e9, "n_rays*n_samples must fit 32 bits; split the rays");
  {
    // tvm_forward accepts a workspace below the worst case: the entry list is then bounded by what fits (at least one
    // entry per ray on average) and the caller checks tvm_forward_entries; everything that reads a stash
    // (tvm_backward*) needs the worst-case size
    const uint32_t cap = workspace_capacity(n, S, ws_bytes);
    const size_t worst = (size_t)n * (size_t)S;
    TVM_REQUIRE(cap == worst || (bounded_ok && cap >= std::min(worst, (size_t)n)),
                "workspace too small: need %zu bytes%s, got %zu", carve_workspace(nullptr, n, S).bytes,
                bounded_ok ? " (or at least room for one entry per ray)" : "", ws_bytes);
    P.ws = carve_workspace(ws, n, S, cap);
  }
  TVM_REQUIRE(P.ws.bytes <= ws_bytes, "workspace too small: need %zu bytes, got %zu", P.ws.bytes, ws_bytes);
  TVM_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  P.NB = P.ws.NB;
  P.in_mlp_c = in_mlp_c(*m);
  P.st = app_tile_stride(m->n_app, P.in_mlp_c);
  P.counters = nullptr;
  P.aux = TvmAux{};
  P.bg_rand = nullptr;
  P.bg = TvmBgNet{};
  return 0;
}

// SM count of the CURRENT device (a process may drive several GPUs through one copy of the library)
static int device_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  int sms = cache[dev].load(std::memory_order_relaxed);
  if (sms <= 0) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cache[dev].store(sms, std::memory_order_relaxed);
  }
  return sms;
}

}  // namespace tvm

using namespace tvm;

extern "C" int tvm_workspace_bytes(int n_rays, int n_samples, size_t* out_bytes) {
  TVM_REQUIRE(out_bytes && n_rays > 0 && n_samples > 0, "bad arguments");
  TVM_REQUIRE((double)n_rays * n_samples < 4.0e9, "n_rays*n_samples must fit 32 bits; split the rays");
  *out_bytes = carve_workspace(nullptr, n_rays, n_samples).bytes;
  return 0;
}

extern "C" int tvm_workspace_layout(int n_rays, int n_samples, TvmWorkspaceLayout* out) {
  TVM_REQUIRE(out && n_rays > 0 && n_samples > 0, "bad arguments");
  TVM_REQUIRE((double)n_rays * n_samples < 4.0e9, "n_rays*n_samples must fit 32 bits; split the rays");
  const Workspace w = carve_workspace((void*)256, n_rays, n_samples);      // non-null base so that members get addresses
  auto off = [](const void* p) { return (size_t)((const char*)p - (const char*)256); };
  out->n_entries = off(w.n_entries);
  out->blk_mask = off(w.blk_mask);
  out->blk_base = off(w.blk_base);
  out->ent = off(w.ent);
  out->ent_w = off(w.ent_w);
  out->ent_rgb = off(w.ent_rgb);
  out->acc = off(w.acc);
  out->rgb_sum = off(w.rgb_sum);
  out->capacity = w.cap;
  out->n_blocks = w.NB;
  out->bytes = w.bytes;
  return 0;
}

extern "C" int tvm_workspace_bytes_bounded(int n_rays, int n_samples, uint32_t max_entries, size_t* out_bytes) {
  TVM_REQUIRE(out_bytes && n_rays > 0 && n_samples > 0, "bad arguments");
  TVM_REQUIRE((double)n_rays * n_samples < 4.0e9, "n_rays*n_samples must fit 32 bits; split the rays");
  const size_t floor_entries = std::min((size_t)n_rays * (size_t)n_samples, (size_t)n_rays);
  *out_bytes = carve_workspace(nullptr, n_rays, n_samples, std::max((size_t)max_entries, floor_entries)).bytes;
  return 0;
}

extern "C" int tvm_workspace_capacity(int n_rays, int n_samples, size_t ws_bytes, uint32_t* out_entries) {
  TVM_REQUIRE(out_entries && n_rays > 0 && n_samples > 0, "bad arguments");
  TVM_REQUIRE((double)n_rays * n_samples < 4.0e9, "n_rays*n_samples must fit 32 bits; split the rays");
  *out_entries = workspace_capacity(n_rays, n_samples, ws_bytes);
  return 0;
}

extern "C" int tvm_forward_entries(const void* ws, void* stream, uint32_t* out_entries) {
  TVM_REQUIRE(ws && out_entries, "null argument");
  TVM_CHECK_CUDA(cudaMemcpyAsync(out_entries, ws, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  TVM_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

namespace tvm {
int launch_bg(const FwdParams& P, int num_sms, cudaStream_t stream);   // tvm_bg.cu
}

// shared body of tvm_forward (bg_host == NULL) and tvm_forward_npp
static int forward_impl(const TvmModel* m_host, const TvmBgNet* bg_host, const float* rays, int n_rays, int n_samples,
                        const float* jitter, const float* bg_rand, uint32_t flags, float* rgb_map, float* depth_map,
                        const TvmAux* aux_host, uint64_t* counters, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FwdParams P;
  if (int rc = fill_fwd_params(P, m_host, rays, n_rays, n_samples, jitter, flags, ws, ws_bytes, true)) return rc;
  TVM_REQUIRE(rgb_map && depth_map, "null output");
  TVM_REQUIRE((m_host->sampling == TVM_SAMPLING_NPP) == (bg_host != nullptr),
              "TVM_SAMPLING_NPP models go through tvm_forward_npp, all others through tvm_forward");
  P.bg_rand = bg_rand;
  if (bg_host) {
    TVM_REQUIRE(jitter && bg_rand && n_samples >= 2, "tvm_forward_npp needs fg_rand [n][S], bg_rand [n][512], S >= 2");
    P.bg = *bg_host;
    P.flags = flags & ~TVM_WHITE_BG;          // the foreground always renders on black (nerfplusplus.py:274)
  }
  P.rgb_map = rgb_map;
  P.depth_map = depth_map;
  P.counters = (unsigned long long*)counters;
  if (aux_host) P.aux = *aux_host;
  // the parity instantiation (no ERT, every mask bit written) is selected by the per-sample outputs only
  const bool has_aux = aux_host && (P.aux.bbox_bits || P.aux.valid_bits || P.aux.app_bits || P.aux.sigma ||
                                    P.aux.weight || P.aux.rgb || P.aux.acc_map);
  const size_t nb_bytes = (size_t)n_rays * P.NB * 4;
  // TVM_EVAL_ONLY: composite inside the appearance head (fixed-point sums per ray in the space of the block tables)
  const bool fused = (flags & TVM_EVAL_ONLY) && !has_aux && !bg_host && P.m.variant == TVM_VARIANT_VM && fix_sums_fit(n_rays, P.NB);
  if (!fused) P.flags &= ~TVM_EVAL_ONLY;
  TVM_CHECK_CUDA(cudaMemsetAsync(P.ws.n_entries, 0, 256, stream));
  if (fused) TVM_CHECK_CUDA(cudaMemsetAsync(fix_sums(P.ws), 0, (size_t)n_rays * 12, stream));
  else TVM_CHECK_CUDA(cudaMemsetAsync(P.ws.blk_mask, 0, nb_bytes, stream));
  if (has_aux) {
    if (P.aux.bbox_bits) TVM_CHECK_CUDA(cudaMemsetAsync(P.aux.bbox_bits, 0, nb_bytes, stream));
    if (P.aux.valid_bits) TVM_CHECK_CUDA(cudaMemsetAsync(P.aux.valid_bits, 0, nb_bytes, stream));
    if (P.aux.app_bits) TVM_CHECK_CUDA(cudaMemsetAsync(P.aux.app_bits, 0, nb_bytes, stream));
    if (P.aux.sigma) TVM_CHECK_CUDA(cudaMemsetAsync(P.aux.sigma, 0, (size_t)n_rays * n_samples * 4, stream));
    if (P.aux.weight) TVM_CHECK_CUDA(cudaMemsetAsync(P.aux.weight, 0, (size_t)n_rays * n_samples * 4, stream));
    if (P.aux.rgb) TVM_CHECK_CUDA(cudaMemsetAsync(P.aux.rgb, 0, (size_t)n_rays * n_samples * 12, stream));
  }
  const int march_blocks = (n_rays + kMarchWarps - 1) / kMarchWarps;
  {
    ProfileScope prof(TVM_STAGE_MARCH, stream);
    const bool npp = P.m.sampling == TVM_SAMPLING_NPP, c16 = P.m.n_density == 16;
    void (*kern)(const FwdParams) =
        has_aux ? (npp ? (c16 ? k_march<true, true, 16> : k_march<true, true, 0>)
                       : (c16 ? k_march<true, false, 16> : k_march<true, false, 0>))
                : (npp ? (c16 ? k_march<false, true, 16> : k_march<false, true, 0>)
                       : (c16 ? k_march<false, false, 16> : k_march<false, false, 0>));
    kern<<<march_blocks, kMarchWarps * 32, 0, stream>>>(P);
  }
  TVM_CHECK_CUDA(cudaGetLastError());

  const uint32_t mlp = flags & TVM_MLP_MASK;
  if (mlp == TVM_MLP_FP32) {
    const size_t smem = ((size_t)kAppTile * 2 * P.st + kAppTile * 8) * sizeof(float);
    auto kern = P.m.variant == TVM_VARIANT_REF ? k_app_simt<TVM_REF_HEAD_LD> : k_app_simt<32>;
    TVM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TVM_REQUIRE(smem <= 200 * 1024, "appearance tile does not fit shared memory");
    {
      ProfileScope prof(TVM_STAGE_APP, stream);
      kern<<<device_sms() * 2, kAppThreads, smem, stream>>>(P);
    }
    TVM_CHECK_CUDA(cudaGetLastError());
  } else {
    ProfileScope prof(TVM_STAGE_APP, stream);
    if (int rc = launch_app_tc(P, device_sms(), stream)) return rc;
  }
  {
    ProfileScope prof(TVM_STAGE_COMPOSITE, stream);
    if (fused) k_finalize<<<(n_rays + 255) / 256, 256, 0, stream>>>(P);
    else k_composite<<<(n_rays + 7) / 8, 256, 0, stream>>>(P);
  }
  TVM_CHECK_CUDA(cudaGetLastError());
  if (bg_host) {
    ProfileScope prof(TVM_STAGE_BG, stream);
    if (int rc = launch_bg(P, device_sms(), stream)) return rc;
  }
  return 0;
}

extern "C" int tvm_forward(const TvmModel* m_host, const float* rays, int n_rays, int n_samples,
                           const float* jitter, uint32_t flags, float* rgb_map, float* depth_map,
                           const TvmAux* aux_host, uint64_t* counters, void* ws, size_t ws_bytes,
                           void* stream_) {
  return forward_impl(m_host, nullptr, rays, n_rays, n_samples, jitter, nullptr, flags, rgb_map, depth_map, aux_host,
                      counters, ws, ws_bytes, stream_);
}

extern "C" int tvm_forward_npp(const TvmModel* m_host, const TvmBgNet* bg_host, const float* rays, int n_rays,
                               int n_samples, const float* fg_rand, const float* bg_rand, uint32_t flags,
                               float* rgb_map, float* depth_map, const TvmAux* aux_host, uint64_t* counters, void* ws,
                               size_t ws_bytes, void* stream_) {
  TVM_REQUIRE(bg_host != nullptr, "null TvmBgNet");
  return forward_impl(m_host, bg_host, rays, n_rays, n_samples, fg_rand, bg_rand, flags, rgb_map, depth_map, aux_host,
                      counters, ws, ws_bytes, stream_);
}
