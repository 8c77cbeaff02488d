// Parameter upload / layout conversion and small utility entry points of libtvmrender.
//
// The reference keeps parameters as Jittor NCHW Vars (tensorf-myc/models/tensoRF.py:154-164); the
// kernels want channels-last grids (one 64 B / 192 B segment per texel), a bit-packed alpha volume
// (the reference itself stores it bit-packed on disk, tensorBase.py:253-264) and [in][out] linear
// weights.  These kernels are pure bandwidth: coalesced on the write side, strided on the read side
// through a 32x33 shared-memory transpose tile.
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tvm_common.cuh"

namespace tvm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Measurement hook (tvm_profile_*): the only process-global state of the library.  One mutex guards the record list and
// the event pool; a launch keeps the index of its own record, so concurrent launches on different streams / threads
// cannot close each other's bracket.  Events are per device (an event may only be recorded on a stream of its device).
struct ProfRec { int stage, dev; cudaEvent_t a, b; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::vector<std::pair<int, cudaEvent_t>> g_ev_pool;
static thread_local std::vector<size_t> g_prof_open;      // indices of this thread's open brackets (they nest at most once)
static cudaEvent_t get_event(int dev) {
  for (size_t i = 0; i < g_ev_pool.size(); ++i)
    if (g_ev_pool[i].first == dev) {
      cudaEvent_t e = g_ev_pool[i].second;
      g_ev_pool.erase(g_ev_pool.begin() + i);
      return e;
    }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
bool profile_on() { return g_prof_on.load(std::memory_order_relaxed); }
void profile_begin(int stage, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{stage, dev, get_event(dev), get_event(dev)};
  cudaEventRecord(r.a, s);
  g_prof_open.push_back(g_prof.size());
  g_prof.push_back(r);
}
void profile_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_open.empty()) return;
  const size_t i = g_prof_open.back();
  g_prof_open.pop_back();
  if (i < g_prof.size()) cudaEventRecord(g_prof[i].b, s);
}

int validate_model(const TvmModel& m) {
  TVM_REQUIRE(m.n_density > 0 && m.n_density % 4 == 0, "n_density must be a positive multiple of 4");
  TVM_REQUIRE(m.n_app > 0 && m.n_app % 4 == 0, "n_app must be a positive multiple of 4");
  TVM_REQUIRE(m.app_dim > 0 && m.app_dim <= kMaxAppDim, "app_dim must be in 1..32");
  TVM_REQUIRE(m.feature_c == kFeatureC, "featureC must be 128");
  TVM_REQUIRE(m.view_pe >= 0 && m.fea_pe >= 0 && m.view_pe <= 8 && m.fea_pe <= 8, "pe out of range");
  for (int i = 0; i < 3; ++i) {
    TVM_REQUIRE(m.grid[i] >= 2, "gridSize must be >= 2");
    TVM_REQUIRE(m.density_plane[i] && m.density_line[i] && m.app_plane[i] && m.app_line[i], "null grid pointer");
  }
  TVM_REQUIRE(m.basis_t && m.w1_t && m.b1 && m.w2_t && m.b2 && m.w3 && m.b3, "null MLP pointer");
  TVM_REQUIRE(m.variant == TVM_VARIANT_VM || m.variant == TVM_VARIANT_REF, "unknown variant");
  if (m.variant == TVM_VARIANT_REF) {
    TVM_REQUIRE(m.head_bias != nullptr, "TVM_VARIANT_REF needs head_bias");
    TVM_REQUIRE(m.app_dim + 8 <= TVM_REF_HEAD_LD, "app_dim + 8 heads must fit 48 columns");
  }
  if (m.alpha_bits) {
    for (int i = 0; i < 3; ++i) TVM_REQUIRE(m.alpha_grid[i] >= 2, "alpha grid must be >= 2");
  }
  return 0;
}

// [C][HW] -> [HW][C]
__global__ void k_transpose(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = by + j, c = bx + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = bx + j, r = by + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

static int transpose(const float* in, float* out, int rows, int cols, cudaStream_t s) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  TVM_REQUIRE(grid.y <= 65535, "too many rows for transpose");
  k_transpose<<<grid, block, 0, s>>>(in, out, rows, cols);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// several transposes in ONE launch (the 12 factor grids of a model: a training step packs them after every optimiser step
// and unpacks their gradients -- 24 launches of a launch-bound step become 2)
struct TransposeBatch {
  TvmTransposeJob job[TVM_TRANSPOSE_MAX];
  int tile_end[TVM_TRANSPOSE_MAX];     // exclusive prefix ends of the jobs' 32x32 tile ranges
  int n;
};
__global__ void __launch_bounds__(256) k_transpose_batch(const TransposeBatch B) {
  __shared__ float tile[32][33];
  int t = blockIdx.x, j = 0;
  while (j < B.n - 1 && t >= B.tile_end[j]) ++j;
  t -= j ? B.tile_end[j - 1] : 0;
  const float* __restrict__ in = B.job[j].src;
  float* __restrict__ out = B.job[j].dst;
  const int rows = B.job[j].rows, cols = B.job[j].cols;
  const int sld = B.job[j].src_ld ? B.job[j].src_ld : cols, dld = B.job[j].dst_ld ? B.job[j].dst_ld : rows;
  const int tiles_x = (cols + 31) / 32;
  const int bx = (t % tiles_x) * 32, by = (t / tiles_x) * 32;
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const int r = by + jj, c = bx + threadIdx.x;
    if (r < rows && c < cols) tile[jj][threadIdx.x] = in[(size_t)r * sld + c];
  }
  __syncthreads();
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const int c = bx + jj, r = by + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * dld + r] = tile[threadIdx.x][jj];
  }
}

// [out][in] -> [in][out_pad] with zero padding
__global__ void k_pack_linear(const float* __restrict__ w, int out_c, int in_c, int out_pad, float* __restrict__ t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in_c * out_pad) return;
  const int j = i / out_pad, o = i % out_pad;
  t[i] = o < out_c ? w[(size_t)o * in_c + j] : 0.0f;
}
__global__ void k_unpack_linear(const float* __restrict__ t, int out_c, int in_c, int out_pad, float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in_c * out_c) return;
  const int o = i / in_c, j = i % in_c;
  w[i] = t[(size_t)j * out_pad + o];
}

// one warp per output word: bit = volume > 0
__global__ void k_pack_alpha(const float* __restrict__ vol, size_t n_vox, uint32_t* __restrict__ bits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < n_vox && vol[i] > 0.0f;
  const uint32_t word = __ballot_sync(0xffffffffu, on);
  if ((threadIdx.x & 31) == 0 && (i >> 5) < (n_vox + 31) / 32) bits[i >> 5] = word;
}

// one thread per 8x8x8 brick: OR of its voxels' bits; one warp writes one output word
__global__ void k_pack_bricks(const uint32_t* __restrict__ bits, int D, int H, int W, int n_bricks,
                              uint32_t* __restrict__ bricks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int BW = (W + 7) >> 3, BH = (H + 7) >> 3;
  bool on = false;
  if (i < n_bricks) {
    const int bx = i % BW, by = (i / BW) % BH, bz = i / (BW * BH);
    for (int z = bz * 8; z < min(D, bz * 8 + 8) && !on; ++z)
      for (int y = by * 8; y < min(H, by * 8 + 8) && !on; ++y)
        for (int x = bx * 8; x < min(W, bx * 8 + 8); ++x) {
          const uint32_t idx = ((uint32_t)z * H + y) * W + x;
          if ((bits[idx >> 5] >> (idx & 31u)) & 1u) { on = true; break; }
        }
  }
  const uint32_t word = __ballot_sync(0xffffffffu, on);
  if ((threadIdx.x & 31) == 0 && (i >> 5) < (n_bricks + 31) / 32) bricks[i >> 5] = word;
}

// Neighbourhood words of the brick index: one uint32 per brick, bit (dz+1)*9 + (dy+1)*3 + (dx+1) = brick (bx+dx, by+dy, bz+dz)
// is non-empty (bricks outside the grid: 0).  A run of samples whose voxel box spans at most 3 bricks per axis lies inside the
// 3x3x3 neighbourhood of its middle brick: the exact "does any brick of the box hold a voxel" is then ONE load and one AND
// with the box's 27-bit mask (bricks_maybe, tvm_math.cuh) instead of a loop over up to 27 bricks.
__global__ void k_pack_bricks3(const uint32_t* __restrict__ bricks, int BD, int BH, int BW, int n_bricks,
                               uint32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bricks) return;
  const int bx = i % BW, by = (i / BW) % BH, bz = i / (BW * BH);
  uint32_t word = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = bx + dx, y = by + dy, z = bz + dz;
        if (x < 0 || y < 0 || z < 0 || x >= BW || y >= BH || z >= BD) continue;
        const uint32_t idx = ((uint32_t)z * BH + y) * BW + x;
        if ((bricks[idx >> 5] >> (idx & 31u)) & 1u) word |= 1u << ((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
      }
  out[i] = word;
}

// TensorBase.compute_alpha on arbitrary points (tensorBase.py:451-473): one thread per point.
__global__ void k_density_alpha(const TvmModel m, const float* __restrict__ xyz, int n, float length,
                                float* __restrict__ alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]};
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
  alpha[i] = 1.0f - expf(-sigma * length);
}

// train.py:228: loss = mean((rgb_map - target)^2); d_rgb_map = grad_scale * 2 (rgb_map - target) / (3n)
__global__ void k_mse(const float* __restrict__ rgb, const float* __restrict__ tgt, int n3, float gscale,
                      float* __restrict__ loss, float* __restrict__ d_rgb) {
  float part = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x) {
    const float d = rgb[i] - tgt[i];
    part += d * d;
    if (d_rgb) d_rgb[i] = gscale * 2.0f * d / (float)n3;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0 && loss) atomicAdd(loss, part / (float)n3);
}

}  // namespace tvm

using namespace tvm;

extern "C" const char* tvm_last_error(void) { return g_err; }
extern "C" int tvm_abi_version(void) { return TVM_ABI_VERSION; }

extern "C" int tvm_profile_enable(int on) {
  g_prof_on.store(on != 0);
  return 0;
}

extern "C" int tvm_profile_collect(float* ms_by_stage, int* launches_by_stage) {
  TVM_REQUIRE(ms_by_stage && launches_by_stage, "bad arguments");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    TVM_CHECK_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.0f;
    TVM_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.stage >= 0 && r.stage < TVM_STAGE_COUNT) {
      ms_by_stage[r.stage] += ms;
      launches_by_stage[r.stage] += 1;
    }
    g_ev_pool.push_back({r.dev, r.a});
    g_ev_pool.push_back({r.dev, r.b});
  }
  g_prof.clear();
  return 0;
}

extern "C" int tvm_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return -2;
  }
  return n;
}

extern "C" int tvm_pack_grid(const float* nchw, int C, int H, int W, float* out_hwc, void* stream) {
  TVM_REQUIRE(nchw && out_hwc && C > 0 && H > 0 && W > 0, "bad arguments");
  return transpose(nchw, out_hwc, C, H * W, (cudaStream_t)stream);
}

extern "C" int tvm_unpack_grid(const float* hwc, int C, int H, int W, float* out_nchw, void* stream) {
  TVM_REQUIRE(hwc && out_nchw && C > 0 && H > 0 && W > 0, "bad arguments");
  return transpose(hwc, out_nchw, H * W, C, (cudaStream_t)stream);
}

extern "C" int tvm_pack_linear(const float* w, int out_c, int in_c, int out_pad, float* out_t, void* stream) {
  TVM_REQUIRE(w && out_t && out_c > 0 && in_c > 0 && out_pad >= out_c, "bad arguments");
  const int n = in_c * out_pad;
  k_pack_linear<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, out_c, in_c, out_pad, out_t);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_unpack_linear(const float* w_t, int out_c, int in_c, int out_pad, float* out_w, void* stream) {
  TVM_REQUIRE(w_t && out_w && out_c > 0 && in_c > 0 && out_pad >= out_c, "bad arguments");
  const int n = in_c * out_c;
  k_unpack_linear<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_t, out_c, in_c, out_pad, out_w);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_pack_alpha(const float* volume, int D, int H, int W, uint32_t* bits, void* stream) {
  TVM_REQUIRE(volume && bits && D > 0 && H > 0 && W > 0, "bad arguments");
  const size_t n = (size_t)D * H * W;
  const size_t n_pad = (n + 31) / 32 * 32;
  k_pack_alpha<<<(unsigned)((n_pad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(volume, n, bits);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dil bit (z,y,x) = OR of the 2x2x2 voxel bits at (x..x+1, y..y+1, z..z+1) (clipped at the upper faces): the alpha-mask
// decision of a sample whose 8 trilinear taps are all in range with non-zero weights, in ONE bit lookup
__global__ void k_pack_dilated(const uint32_t* __restrict__ bits, int D, int H, int W, uint32_t* __restrict__ dil) {
  const size_t total = (size_t)D * H * W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool set = false;
  if (i < total) {
    const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((size_t)W * H));
    for (int dz = 0; dz < 2 && !set; ++dz)
      for (int dy = 0; dy < 2 && !set; ++dy)
        for (int dx = 0; dx < 2 && !set; ++dx) {
          const int xx = x + dx, yy = y + dy, zz = z + dz;
          if (xx < W && yy < H && zz < D) {
            const size_t j = ((size_t)zz * H + yy) * W + xx;
            set = (bits[j >> 5] >> (j & 31)) & 1u;
          }
        }
  }
  const uint32_t word = __ballot_sync(0xffffffffu, set);
  if ((threadIdx.x & 31) == 0 && (i >> 5) < (total + 31) / 32) dil[i >> 5] = word;
}

extern "C" int tvm_pack_alpha_dilated(const uint32_t* bits, int D, int H, int W, uint32_t* dilated, void* stream) {
  TVM_REQUIRE(bits && dilated && D > 0 && H > 0 && W > 0, "bad arguments");
  const size_t warps = ((size_t)D * H * W + 31) / 32;
  k_pack_dilated<<<(unsigned)((warps + 3) / 4), 128, 0, (cudaStream_t)stream>>>(bits, D, H, W, dilated);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_pack_alpha_bricks(const uint32_t* bits, int D, int H, int W, uint32_t* bricks, void* stream) {
  TVM_REQUIRE(bits && bricks && D > 0 && H > 0 && W > 0, "bad arguments");
  const int n = ((D + 7) / 8) * ((H + 7) / 8) * ((W + 7) / 8);
  const int n_pad = (n + 31) / 32 * 32;
  k_pack_bricks<<<(n_pad + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bits, D, H, W, n, bricks);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_pack_alpha_bricks3(const uint32_t* bricks, int D, int H, int W, uint32_t* bricks3, void* stream) {
  TVM_REQUIRE(bricks && bricks3 && D > 0 && H > 0 && W > 0, "bad arguments");
  const int BD = (D + 7) / 8, BH = (H + 7) / 8, BW = (W + 7) / 8;
  const int n = BD * BH * BW;
  k_pack_bricks3<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bricks, BD, BH, BW, n, bricks3);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_density_alpha(const TvmModel* m_host, const float* xyz, int n_pts, float length,
                                 float* alpha_out, void* stream) {
  TVM_REQUIRE(m_host && xyz && alpha_out && n_pts > 0, "bad arguments");
  if (int rc = validate_model(*m_host)) return rc;
  k_density_alpha<<<(n_pts + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*m_host, xyz, n_pts, length, alpha_out);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_mse_loss(const float* rgb_map, const float* target, int n_rays, float grad_scale,
                            float* loss_out, float* d_rgb_map, void* stream) {
  TVM_REQUIRE(rgb_map && target && n_rays > 0, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (loss_out) TVM_CHECK_CUDA(cudaMemsetAsync(loss_out, 0, 4, s));
  const int n3 = n_rays * 3;
  int blocks = (n3 + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  k_mse<<<blocks, 256, 0, s>>>(rgb_map, target, n3, grad_scale, loss_out, d_rgb_map);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Pair records of a channels-last grid (TvmModel.app_plane_pair / app_line_pair): one thread per 16-byte group
// [t0 c..c+3 | t1 c..c+3] with t0 = (r, x), t1 = (r, min(x+1, W-1)).
__global__ void k_pack_pair16(const float* __restrict__ src, size_t n_groups, int W, int C4, uint4* __restrict__ dst, int h16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_groups; i += (size_t)gridDim.x * blockDim.x) {
    const size_t texel = i / (size_t)C4;
    const int g = (int)(i - texel * (size_t)C4);
    const int x = (int)(texel % (size_t)W);
    const size_t nb = (x + 1 < W) ? texel + 1 : texel;
    const float4 a = *reinterpret_cast<const float4*>(src + (texel * (size_t)C4 + g) * 4);
    const float4 b = *reinterpret_cast<const float4*>(src + (nb * (size_t)C4 + g) * 4);
    uint4 o;
    if (h16) {
      const __half2 a0 = __floats2half2_rn(a.x, a.y), a1 = __floats2half2_rn(a.z, a.w);
      const __half2 b0 = __floats2half2_rn(b.x, b.y), b1 = __floats2half2_rn(b.z, b.w);
      o = make_uint4(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1),
                     *reinterpret_cast<const uint32_t*>(&b0), *reinterpret_cast<const uint32_t*>(&b1));
    } else {
      const __nv_bfloat162 a0 = __floats2bfloat162_rn(a.x, a.y), a1 = __floats2bfloat162_rn(a.z, a.w);
      const __nv_bfloat162 b0 = __floats2bfloat162_rn(b.x, b.y), b1 = __floats2bfloat162_rn(b.z, b.w);
      o = make_uint4(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1),
                     *reinterpret_cast<const uint32_t*>(&b0), *reinterpret_cast<const uint32_t*>(&b1));
    }
    dst[i] = o;
  }
}

extern "C" int tvm_pack_pair16(const float* src, int rows, int W, int C, void* dst, uint32_t flags, void* stream) {
  const uint32_t mode = flags & TVM_MLP_MASK;
  TVM_REQUIRE(src && dst && rows > 0 && W > 0 && C > 0 && C % 4 == 0 && (mode == TVM_MLP_BF16 || mode == TVM_MLP_FP16),
              "bad arguments");
  TVM_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0,
              "tvm_pack_pair16: src and dst must be 16-byte aligned");
  const size_t n_groups = (size_t)rows * (size_t)W * (size_t)(C / 4);
  const int blocks = (int)std::min<size_t>((n_groups + 255) / 256, 148 * 16);
  k_pack_pair16<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, n_groups, W, C / 4, (uint4*)dst, mode == TVM_MLP_FP16);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_transpose_batch(const TvmTransposeJob* jobs_host, int n_jobs, void* stream) {
  TVM_REQUIRE(jobs_host && n_jobs > 0, "bad arguments");
  for (int base = 0; base < n_jobs; base += TVM_TRANSPOSE_MAX) {
    TransposeBatch B;
    B.n = std::min(TVM_TRANSPOSE_MAX, n_jobs - base);
    long long tiles = 0;
    for (int i = 0; i < B.n; ++i) {
      const TvmTransposeJob& j = jobs_host[base + i];
      TVM_REQUIRE(j.src && j.dst && j.rows > 0 && j.cols > 0 && (j.src_ld == 0 || j.src_ld >= j.cols) &&
                  (j.dst_ld == 0 || j.dst_ld >= j.rows), "bad transpose job %d", base + i);
      B.job[i] = j;
      tiles += (long long)((j.rows + 31) / 32) * ((j.cols + 31) / 32);
      TVM_REQUIRE(tiles < (1ll << 31), "too many tiles");
      B.tile_end[i] = (int)tiles;
    }
    k_transpose_batch<<<(unsigned)tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(B);
    TVM_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
