// The callers either side of the ray path, references relative to tensorf-myc/:
//
//   k_dense_alpha      getDenseAlpha (models/tensorBase.py:366-384): compute_alpha on the align_corners lattice,
//                      written transposed [z][y][x] and clamped as updateAlphaMask does (:390-391)
//   k_alpha_pool_pack  updateAlphaMask (:386-409): 3x3x3 max pool, >= threshold, bit-pack (+ optional {0,1} fp32
//                      volume), index bounding box of the surviving voxels, voxel count
//   k_filter_rays      filtering_rays (:411-441): bbox-only slab test, or "any of N_samples hits the alpha mask"
//   k_generate_rays    get_ray_directions(+_blender) / get_rays (dataLoader/ray_utils.py:81-153) with the
//                      normalisation of dataLoader/blender.py:75
//   k_upsample         up_sampling_VM (models/tensoRF.py:248-262): bilinear, align_corners=True, NCHW -> NCHW
#include "tvm_common.cuh"

namespace tvm {

// torch.linspace(0, 1, n)[i] in fp32: evaluated from both ends (assumption A8 of oracle/maintain_oracle.py)
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n <= 1) return 0.0f;
  const float step = TVM_DIV(1.0f, (float)(n - 1));
  return i < n / 2 ? TVM_MUL(step, (float)i) : TVM_SUB(1.0f, TVM_MUL(step, (float)(n - 1 - i)));
}

__device__ __forceinline__ float density_alpha_at(const TvmModel& m, const float p[3], float length) {
  bool ok = true;
  if (m.alpha_bits) ok = alpha_mask_test(m, m.alpha_bits, p);
  float sigma = 0.0f;
  if (ok) {
    float u[3];
    grid_coords(m, p, u);
    Axis ax[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) ax[a] = axis_taps(u[a], m.grid[a]);
    float f = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const VmTaps t = vm_taps(m, ax, k);
      for (int c = 0; c < m.n_density; c += 4) {
        float4 pv, lv;
        vm_sample4(m.density_plane[k], m.density_line[k], t, m.n_density, c, pv, lv);
        f += pv.x * lv.x + pv.y * lv.y + pv.z * lv.z + pv.w * lv.w;
      }
    }
    sigma = feature2density(m, f);
  }
  return 1.0f - expf(-sigma * length);
}

__global__ void __launch_bounds__(256) k_dense_alpha(const TvmModel m, int gx, int gy, int gz, float length,
                                                     float* __restrict__ alpha_zyx) {
  const size_t total = (size_t)gx * gy * gz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % gx), y = (int)((i / gx) % gy), z = (int)(i / ((size_t)gx * gy));
    const float s[3] = {linspace01(x, gx), linspace01(y, gy), linspace01(z, gz)};
    float p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)      // dense_xyz = aabb[0] * (1 - samples) + aabb[1] * samples   (:376)
      p[a] = TVM_ADD(TVM_MUL(m.aabb[a], TVM_SUB(1.0f, s[a])), TVM_MUL(m.aabb[3 + a], s[a]));
    const float al = density_alpha_at(m, p, length);
    alpha_zyx[i] = fminf(fmaxf(al, 0.0f), 1.0f);
  }
}

// out[6] = {min x, min y, min z, max x, max y, max z} (voxel indices), n_set += number of set voxels
__global__ void __launch_bounds__(256) k_alpha_pool_pack(const float* __restrict__ alpha, int gx, int gy, int gz,
                                                         float thres, float* __restrict__ volume,
                                                         uint32_t* __restrict__ bits, int* __restrict__ bbox,
                                                         unsigned long long* __restrict__ n_set) {
  const size_t total = (size_t)gx * gy * gz;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // grid covers ceil(total/32) warps exactly
  bool set = false;
  int x = 0, y = 0, z = 0;
  if (i < total) {
    x = (int)(i % gx);
    y = (int)((i / gx) % gy);
    z = (int)(i / ((size_t)gx * gy));
    float mx = -INFINITY;                    // max_pool3d pads with -inf (A10)
    for (int dz = -1; dz <= 1; ++dz) {
      const int zz = z + dz;
      if (zz < 0 || zz >= gz) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= gy) continue;
        const float* row = alpha + ((size_t)zz * gy + yy) * gx;
        mx = fmaxf(mx, row[x]);
        if (x > 0) mx = fmaxf(mx, row[x - 1]);
        if (x + 1 < gx) mx = fmaxf(mx, row[x + 1]);
      }
    }
    set = mx >= thres;
    if (volume) volume[i] = set ? 1.0f : 0.0f;
  }
  const uint32_t word = __ballot_sync(0xffffffffu, set);
  const int lane = threadIdx.x & 31;
  if (lane == 0 && (i >> 5) < (total + 31) / 32) bits[i >> 5] = word;
  if (word == 0) return;
  const int big = 0x7fffffff;
  const int mnx = __reduce_min_sync(0xffffffffu, set ? x : big), mny = __reduce_min_sync(0xffffffffu, set ? y : big),
            mnz = __reduce_min_sync(0xffffffffu, set ? z : big);
  const int mxx = __reduce_max_sync(0xffffffffu, set ? x : -1), mxy = __reduce_max_sync(0xffffffffu, set ? y : -1),
            mxz = __reduce_max_sync(0xffffffffu, set ? z : -1);
  if (lane == 0) {
    atomicMin(&bbox[0], mnx); atomicMin(&bbox[1], mny); atomicMin(&bbox[2], mnz);
    atomicMax(&bbox[3], mxx); atomicMax(&bbox[4], mxy); atomicMax(&bbox[5], mxz);
    atomicAdd(n_set, (unsigned long long)__popc(word));
  }
}

__global__ void k_init_bbox(int* bbox, unsigned long long* n_set) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0x7fffffff;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = -1;
  else if (threadIdx.x == 6) *n_set = 0ull;
}

// ---- filtering_rays ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_filter_bbox(const TvmModel m, const float* __restrict__ rays, int n,
                                                     uint8_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = rays + 6 * (size_t)i;
  float t_min = -INFINITY, t_max = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float d = r[3 + a], vec = (d == 0.0f) ? 1e-6f : d;
    const float ra = TVM_DIV(TVM_SUB(m.aabb[3 + a], r[a]), vec), rb = TVM_DIV(TVM_SUB(m.aabb[a], r[a]), vec);
    t_min = fmaxf(t_min, fminf(ra, rb));
    t_max = fminf(t_max, fmaxf(ra, rb));
  }
  mask[i] = t_max > t_min ? 1 : 0;
}

// one warp per ray; the reference evaluates sample_alpha on EVERY sample (no bbox gate): zeros padding decides
__global__ void __launch_bounds__(256) k_filter_alpha(const TvmModel m, const float* __restrict__ rays, int n, int S,
                                                      const float* __restrict__ jitter, uint8_t* __restrict__ mask) {
  const int ray = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ray >= n) return;
  RayMarch r;
  ray_setup(m, rays + 6 * (size_t)ray, jitter, ray, S, r);      // jitter: NULL (uniform, is_train=False) or [n][S] (NeRF++)
  bool any = false;
  const int NB = (S + 31) / 32;
  // coarse pass as in k_march: lane l decides block win + l (voxel footprint of the block against the brick index; one
  // lookup in the 3x3x3-dilated index where the neighbourhood is empty), then only the surviving blocks are sampled.
  // No bbox gate (the reference has none here): bricks_maybe alone is conservative for sample_alpha > 0.
  for (int win = 0; win < NB && !any; win += 32) {
    uint32_t visit = 0xffffffffu;
    if (m.alpha_bricks) {
      const int b = win + lane;
      bool maybe = false;
      if (b < NB) {
        const int k0 = b * 32, k1 = min(b * 32 + 31, S - 1);
        float p0[3], p1[3];
        sample_point(m, r, sample_z(m, r, k0), p0);
        sample_point(m, r, sample_z(m, r, k1), p1);
        maybe = bricks_maybe(m, m.alpha_bricks, p0, p1);
      }
      visit = __ballot_sync(0xffffffffu, maybe);
    } else if (NB - win < 32) {
      visit = (1u << (NB - win)) - 1u;
    }
    while (visit && !any) {
      const int b = win + __ffs(visit) - 1;
      visit &= visit - 1;
      const int k = b * 32 + lane;
      float p[3];
      sample_point(m, r, sample_z(m, r, k), p);
      const bool hit = k < S && alpha_mask_test(m, m.alpha_bits, p);
      any = __any_sync(0xffffffffu, hit);
    }
  }
  if (lane == 0) mask[ray] = any ? 1 : 0;
}

// ---- ray generation ------------------------------------------------------------------------------------------------------
struct RayGen {
  float c2w[12];
  float fx, fy, cx, cy;
  int H, W, blender, normalize;
};
__global__ void __launch_bounds__(256) k_generate_rays(const RayGen g, float* __restrict__ rays) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.H * g.W) return;
  const float i = (float)(idx % g.W) + 0.5f, j = (float)(idx / g.W) + 0.5f;
  float d[3];
  d[0] = -TVM_DIV(TVM_SUB(i, g.cx), g.fx);
  d[1] = g.blender ? -TVM_DIV(TVM_SUB(j, g.cy), g.fy) : TVM_DIV(TVM_SUB(j, g.cy), g.fy);
  d[2] = g.blender ? 1.0f : -1.0f;
  if (g.normalize) {
    const float nrm = sqrtf(dot3_seq(d, d));
    d[0] = TVM_DIV(d[0], nrm); d[1] = TVM_DIV(d[1], nrm); d[2] = TVM_DIV(d[2], nrm);
  }
  float* o = rays + 6 * (size_t)idx;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    o[a] = g.c2w[a * 4 + 3];
    o[3 + a] = fmaf(d[2], g.c2w[a * 4 + 2], fmaf(d[1], g.c2w[a * 4 + 1], d[0] * g.c2w[a * 4 + 0]));
  }
}

// ---- bilinear upsample, align_corners = True ---------------------------------------------------------------------------------
// ONE launch for every grid of a model half (three planes + three lines): a job table in the kernel parameters, each job a
// contiguous range of blocks.  A thread owns four consecutive outputs of a row (one 16-byte store) when the row length allows,
// else one output; the flattened (channel, row, quad) index is split with two 32-bit divisions per thread, so every block
// is full whatever the row length (the per-grid kernel it replaces ran 75 of 128 lanes on a 300-wide row and took six
// launches: 16 % of the HBM peak).
struct UpsampleJob {
  const float* src;
  float* dst;
  int C, H, W, H2, W2;
  int vec;                  // 1: W2 % 4 == 0 and dst 16-byte aligned -> float4 stores
  uint32_t items;           // C * H2 * (vec ? W2 / 4 : W2)
  uint32_t block0;          // first block of this job
};
struct UpsampleBatch {
  UpsampleJob job[TVM_UPSAMPLE_MAX_GRIDS];
  int n;
};
__global__ void __launch_bounds__(256) k_upsample_batch(const __grid_constant__ UpsampleBatch B) {
  int j = 0;
#pragma unroll 1
  while (j + 1 < B.n && blockIdx.x >= B.job[j + 1].block0) ++j;
  const UpsampleJob& J = B.job[j];
  const uint32_t item = (blockIdx.x - J.block0) * 256u + threadIdx.x;
  if (item >= J.items) return;
  const int H = J.H, W = J.W, H2 = J.H2, W2 = J.W2;
  const uint32_t per_row = J.vec ? (uint32_t)W2 / 4u : (uint32_t)W2;
  const uint32_t rowi = item / per_row, xq = item - rowi * per_row;       // rowi = c * H2 + y2
  const uint32_t c = rowi / (uint32_t)H2, y2 = rowi - c * (uint32_t)H2;
  const float sh = H2 > 1 ? (float)(H - 1) / (float)(H2 - 1) : 0.0f;
  const float sw = W2 > 1 ? (float)(W - 1) / (float)(W2 - 1) : 0.0f;
  const float fy = sh * (float)y2;
  const int y0 = (int)fy;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, ly0 = 1.0f - ly1;
  const float* __restrict__ r0 = J.src + ((size_t)c * H + y0) * W;
  const float* __restrict__ r1 = J.src + ((size_t)c * H + y1) * W;
  float* __restrict__ out = J.dst + (size_t)rowi * W2;
  auto at = [&](int x2) {
    const float fx = sw * (float)x2;
    const int x0 = (int)fx;
    const int x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float lx1 = fx - (float)x0, lx0 = 1.0f - lx1;
    const float top = lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1);
    const float bot = lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1);
    return ly0 * top + ly1 * bot;
  };
  if (J.vec) {
    const int x2 = (int)xq * 4;
    __stcs(reinterpret_cast<float4*>(out + x2), make_float4(at(x2), at(x2 + 1), at(x2 + 2), at(x2 + 3)));
  } else {
    out[xq] = at((int)xq);
  }
}

static int grid_for(size_t total, int block = 256) {
  size_t b = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace tvm

using namespace tvm;

extern "C" int tvm_dense_alpha(const TvmModel* m_host, const int32_t* grid_host, float length, float* alpha_zyx,
                               void* stream) {
  TVM_REQUIRE(m_host && grid_host && alpha_zyx, "null argument");
  if (int rc = validate_model(*m_host)) return rc;
  const int gx = grid_host[0], gy = grid_host[1], gz = grid_host[2];
  TVM_REQUIRE(gx > 0 && gy > 0 && gz > 0, "bad lattice size");
  k_dense_alpha<<<grid_for((size_t)gx * gy * gz), 256, 0, (cudaStream_t)stream>>>(*m_host, gx, gy, gz, length, alpha_zyx);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_alpha_mask_from_dense(const float* alpha_zyx, const int32_t* grid_host, float thres, float* volume_out,
                                         uint32_t* bits_out, int32_t* bbox_idx, uint64_t* n_set, void* stream) {
  TVM_REQUIRE(alpha_zyx && grid_host && bits_out && bbox_idx && n_set, "null argument");
  const int gx = grid_host[0], gy = grid_host[1], gz = grid_host[2];
  TVM_REQUIRE(gx > 0 && gy > 0 && gz > 0, "bad lattice size");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t total = (size_t)gx * gy * gz;
  k_init_bbox<<<1, 32, 0, s>>>(bbox_idx, (unsigned long long*)n_set);
  const size_t warps = (total + 31) / 32;
  k_alpha_pool_pack<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(alpha_zyx, gx, gy, gz, thres, volume_out, bits_out, bbox_idx,
                                                                (unsigned long long*)n_set);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_filter_rays(const TvmModel* m_host, const float* rays, int n_rays, int n_samples, int bbox_only,
                               const float* fg_rand, uint8_t* mask_out, void* stream) {
  TVM_REQUIRE(m_host && rays && mask_out && n_rays > 0, "bad arguments");
  if (int rc = validate_model(*m_host)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (bbox_only) {
    k_filter_bbox<<<(n_rays + 255) / 256, 256, 0, s>>>(*m_host, rays, n_rays, mask_out);
  } else {
    TVM_REQUIRE(m_host->alpha_bits != nullptr, "filtering_rays(bbox_only=False) needs an alpha mask");
    const bool npp = m_host->sampling == TVM_SAMPLING_NPP;
    TVM_REQUIRE(n_samples > (npp ? 1 : 0), "bad n_samples");
    TVM_REQUIRE(!npp || fg_rand, "TVM_SAMPLING_NPP: filtering_rays samples with NerfPlusPlus.sample_ray and needs fg_rand [n][n_samples]");
    k_filter_alpha<<<(n_rays + 7) / 8, 256, 0, s>>>(*m_host, rays, n_rays, n_samples, npp ? fg_rand : nullptr, mask_out);
  }
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_generate_rays(const float* c2w_host, int H, int W, float fx, float fy, float cx, float cy, int blender,
                                 int normalize, float* rays_out, void* stream) {
  TVM_REQUIRE(c2w_host && rays_out && H > 0 && W > 0, "bad arguments");
  RayGen g;
  for (int i = 0; i < 12; ++i) g.c2w[i] = c2w_host[i];
  g.fx = fx; g.fy = fy; g.cx = cx; g.cy = cy;
  g.H = H; g.W = W; g.blender = blender; g.normalize = normalize;
  k_generate_rays<<<(H * W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, rays_out);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_upsample_grids(int n_grids, const float* const* src_nchw, const int32_t* src_chw, float* const* dst_nchw,
                                  const int32_t* dst_hw, void* stream) {
  TVM_REQUIRE(n_grids > 0 && n_grids <= TVM_UPSAMPLE_MAX_GRIDS && src_nchw && src_chw && dst_nchw && dst_hw, "bad arguments");
  UpsampleBatch B;
  uint32_t blocks = 0;
  for (int i = 0; i < n_grids; ++i) {
    UpsampleJob& J = B.job[i];
    J.src = src_nchw[i];
    J.dst = dst_nchw[i];
    J.C = src_chw[3 * i]; J.H = src_chw[3 * i + 1]; J.W = src_chw[3 * i + 2];
    J.H2 = dst_hw[2 * i]; J.W2 = dst_hw[2 * i + 1];
    TVM_REQUIRE(J.src && J.dst && J.C > 0 && J.H > 0 && J.W > 0 && J.H2 > 0 && J.W2 > 0, "upsample: bad grid %d", i);
    J.vec = ((J.W2 & 3) == 0 && (((uintptr_t)J.dst) & 15) == 0) ? 1 : 0;
    const double items = (double)J.C * J.H2 * (J.vec ? J.W2 / 4 : J.W2);
    TVM_REQUIRE(items < 4.0e9, "upsample: grid %d too large", i);
    J.items = (uint32_t)items;
    J.block0 = blocks;
    blocks += (J.items + 255u) / 256u;
  }
  B.n = n_grids;
  k_upsample_batch<<<blocks, 256, 0, (cudaStream_t)stream>>>(B);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_upsample_grid(const float* src_nchw, int C, int H, int W, float* dst_nchw, int H2, int W2, void* stream) {
  const int32_t chw[3] = {C, H, W}, hw[2] = {H2, W2};
  return tvm_upsample_grids(1, &src_nchw, chw, &dst_nchw, hw, stream);
}

// ---- measurement: L2 gather peak ("report against a measured L2 gather peak") ---------------------------------
// Pure 64-byte gathers with the access shape of k_march's density taps (4 lanes x float4 per tap, 18 independent taps in
// flight per lane group) at pseudo-random texel indices of a buffer that fits the 126 MB L2; no arithmetic beyond a checksum.
namespace tvm {
__global__ void __launch_bounds__(256) k_gather_peak(const float4* __restrict__ buf, uint32_t n_texels, int iters,
                                                     float* __restrict__ sink) {
  const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, q = threadIdx.x & 3;
  uint32_t state = group * 747796405u + 2891336453u;
  float acc = 0.0f;
  for (int it = 0; it < iters; ++it) {
    float4 v[18];
#pragma unroll
    for (int t = 0; t < 18; ++t) {
      state = state * 1664525u + 1013904223u;
      const uint32_t texel = (uint32_t)(((uint64_t)(state >> 4) * n_texels) >> 28);
      v[t] = __ldg(buf + (size_t)texel * 4 + q);
    }
#pragma unroll
    for (int t = 0; t < 18; ++t) acc += v[t].x + v[t].y + v[t].z + v[t].w;
  }
  if (acc == 123.456f) sink[0] = acc;
}
}  // namespace tvm

extern "C" int tvm_bench_gather(const float* buf, size_t n_floats, int n_groups, int iters, float* sink, void* stream) {
  TVM_REQUIRE(buf && sink && n_floats >= 16 && n_groups > 0 && iters > 0, "bad arguments");
  const uint32_t n_texels = (uint32_t)(n_floats / 16);
  const int threads = n_groups * 4;
  k_gather_peak<<<(threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)buf, n_texels, iters, sink);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
