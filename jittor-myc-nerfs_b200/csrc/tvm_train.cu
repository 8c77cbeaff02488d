// Training-step remainder: the whole-grid passes that follow the ray path in train.py:233-261.
// Each regulariser kernel produces the loss value AND adds its gradient in one sweep over the NCHW parameter, so the
// step reads every grid once instead of Jittor's forward + backward graph.  References relative to tensorf-myc/:
//
//   k_tv_loss        utils.TVLoss.execute (utils.py:123-142) as used by TV_loss_density / TV_loss_app (models/tensoRF.py:197-207)
//   k_l1_loss        density_L1 (models/tensoRF.py:191-195): mean |x|
//   k_vector_diffs   vectorDiffs (models/tensoRF.py:177-186): mean |off-diagonal of V V^T|
//   k_adam_multi     jt.optim.Adam(betas=(0.9, 0.99)).step over every parameter tensor in ONE launch (train.py:187,260-261;
//                    update rule: assumption A11 of oracle/maintain_oracle.py)
#include <algorithm>
#include "tvm_common.cuh"

namespace tvm {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.0f;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;       // valid in warp 0
}

// One TV sweep over x [C][H][W]: returns this thread's share of scale_h * sum dh^2 + scale_w * sum dw^2 and adds (or, with
// `overwrite`, stores) d loss / d x.  HBM-bound: per element one read of x, one read-modify-write (or one write) of grad;
// the four neighbours come from L1/L2.  W % 4 == 0 (and 16-byte aligned planes): one thread per four consecutive texels of a
// row, float4 loads / stores, 32-bit index arithmetic (one division per four texels instead of two 64-bit ones per texel).
__device__ __forceinline__ float tv_sweep(const float* __restrict__ x, float* __restrict__ grad, int C, int H, int W,
                                          float scale_h, float scale_w, bool overwrite, unsigned first, unsigned stride) {
  float part = 0.0f;
  const bool vec = (W & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)grad)) & 15) == 0 && (size_t)C * H * W < (1ull << 31);
  if (vec) {
    const unsigned W4 = (unsigned)W >> 2, rows = (unsigned)C * (unsigned)H, n4 = rows * W4;
    for (unsigned g = first; g < n4; g += stride) {
      const unsigned r = g / W4, q = g - r * W4, yy = r % (unsigned)H;
      const float* row = x + (size_t)r * W + 4 * q;
      const float4 c = *reinterpret_cast<const float4*>(row);
      const float4 d = yy + 1 < (unsigned)H ? *reinterpret_cast<const float4*>(row + W) : c;      // next row (diff 0 on the last)
      const float nx = q + 1 < W4 ? row[4] : c.w;                                                 // next texel (diff 0 on the last)
      const float dn[4] = {d.x - c.x, d.y - c.y, d.z - c.z, d.w - c.w};
      const float rt[4] = {c.y - c.x, c.z - c.y, c.w - c.z, nx - c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) part += scale_h * dn[i] * dn[i] + scale_w * rt[i] * rt[i];
      if (grad) {
        const float4 u = yy > 0 ? *reinterpret_cast<const float4*>(row - W) : c;                  // previous row (diff 0 on the first)
        const float pv = q > 0 ? row[-1] : c.x;
        const float up[4] = {c.x - u.x, c.y - u.y, c.z - u.z, c.w - u.w};
        const float lf[4] = {c.x - pv, rt[0], rt[1], rt[2]};
        float4* gp = reinterpret_cast<float4*>(grad + (size_t)r * W + 4 * q);
        float4 o = overwrite ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : *gp;
        o.x += 2.0f * (scale_h * (up[0] - dn[0]) + scale_w * (lf[0] - rt[0]));
        o.y += 2.0f * (scale_h * (up[1] - dn[1]) + scale_w * (lf[1] - rt[1]));
        o.z += 2.0f * (scale_h * (up[2] - dn[2]) + scale_w * (lf[2] - rt[2]));
        o.w += 2.0f * (scale_h * (up[3] - dn[3]) + scale_w * (lf[3] - rt[3]));
        *gp = o;
      }
    }
    return part;
  }
  const size_t total = (size_t)C * H * W;
  for (size_t i = first; i < total; i += stride) {
    const int xx = (int)(i % W), yy = (int)((i / W) % H);
    const float c = x[i];
    const float dn = yy + 1 < H ? x[i + W] - c : 0.0f;       // x[y+1] - x[y]
    const float rt = xx + 1 < W ? x[i + 1] - c : 0.0f;       // x[x+1] - x[x]
    part += scale_h * dn * dn + scale_w * rt * rt;
    if (grad) {
      const float up = yy > 0 ? c - x[i - W] : 0.0f;
      const float lf = xx > 0 ? c - x[i - 1] : 0.0f;
      const float g = 2.0f * (scale_h * (up - dn) + scale_w * (lf - rt));
      grad[i] = overwrite ? g : grad[i] + g;
    }
  }
  return part;
}

// x [C][H][W]; loss += scale_h * sum dh^2 + scale_w * sum dw^2; grad += d loss / d x
__global__ void __launch_bounds__(256) k_tv_loss(const float* __restrict__ x, int C, int H, int W, float scale_h,
                                                 float scale_w, const float* __restrict__ wdev, float* __restrict__ loss,
                                                 float* __restrict__ grad) {
  __shared__ float red[8];
  if (wdev) { scale_h *= *wdev; scale_w *= *wdev; }      // per-step weight of a replayed CUDA graph
  const float part = tv_sweep(x, grad, C, H, W, scale_h, scale_w, false, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
  const float t = block_sum(part, red);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, t);
}

// the TV sweeps of several planes in one launch: blockIdx.y = job
struct TvBatch {
  TvmTvJob job[TVM_TV_MAX];
  float scale_h[TVM_TV_MAX], scale_w[TVM_TV_MAX];
};
__global__ void __launch_bounds__(256) k_tv_loss_batch(const TvBatch B, float* __restrict__ loss) {
  __shared__ float red[8];
  const TvmTvJob& j = B.job[blockIdx.y];
  float scale_h = B.scale_h[blockIdx.y], scale_w = B.scale_w[blockIdx.y];
  if (j.weight_dev) { scale_h *= *j.weight_dev; scale_w *= *j.weight_dev; }
  const float part = tv_sweep(j.plane_nchw, j.grad_nchw, j.C, j.H, j.W, scale_h, scale_w, j.overwrite != 0,
                              blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
  const float t = block_sum(part, red);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, t);
}

__global__ void __launch_bounds__(256) k_l1_loss(const float* __restrict__ x, size_t n, float scale,
                                                 const float* __restrict__ wdev, float* __restrict__ loss,
                                                 float* __restrict__ grad) {
  __shared__ float red[8];
  if (wdev) scale *= *wdev;
  float part = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    part += fabsf(v);
    if (grad) grad[i] += v > 0.0f ? scale : (v < 0.0f ? -scale : 0.0f);
  }
  const float t = block_sum(part, red);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, scale * t);
}

// v [C][L] (C <= 64); one CTA.  loss += scale * sum_{i != j} |<v_i, v_j>|, grad_i += 2 scale sum_{j != i} sign(<v_i, v_j>) v_j
__global__ void __launch_bounds__(256) k_vector_diffs(const float* __restrict__ v, int C, int L, float scale,
                                                      const float* __restrict__ wdev, float* __restrict__ loss,
                                                      float* __restrict__ grad) {
  __shared__ float sgn[64 * 64];
  __shared__ float red[8];
  if (wdev) scale *= *wdev;
  float part = 0.0f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pr = warp; pr < C * C; pr += 8) {         // one warp per (i, j) dot product
    const int i = pr / C, j = pr % C;
    float d = 0.0f;
    for (int l = lane; l < L; l += 32) d = fmaf(v[(size_t)i * L + l], v[(size_t)j * L + l], d);
    d = warp_sum(d);
    if (lane == 0) {
      sgn[pr] = (i == j) ? 0.0f : (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f));
      if (i != j) part += fabsf(d);
    }
  }
  const float t = block_sum(part, red);          // includes the __syncthreads that publishes sgn
  if (threadIdx.x == 0 && loss) atomicAdd(loss, scale * t);
  if (!grad) return;
  for (int e = threadIdx.x; e < C * L; e += blockDim.x) {
    const int i = e / L, l = e % L;
    float g = 0.0f;
    for (int j = 0; j < C; ++j) g = fmaf(sgn[i * C + j], v[(size_t)j * L + l], g);
    grad[e] += 2.0f * scale * g;
  }
}

// ---- Adam over up to TVM_ADAM_MAX_TENSORS tensors in one launch --------------------------------------------------------------
struct AdamTable {
  TvmAdamTensor t[TVM_ADAM_MAX_TENSORS];
  unsigned first_block[TVM_ADAM_MAX_TENSORS + 1];      // blocks [first_block[k], first_block[k+1]) belong to tensor k
  int n;
  float b0, b1, eps, c_step;                           // c_step = sqrt(1 - b1^n) / (1 - b0^n)
  const float* hyper;                                  // device {c_step, lr[...]} overriding the host values (graph replay)
};
constexpr int kAdamPerBlock = 256 * 16;

__global__ void __launch_bounds__(256) k_adam_multi(const AdamTable T) {
  int k = 0;
  while (k + 1 < T.n && blockIdx.x >= T.first_block[k + 1]) ++k;
  const TvmAdamTensor t = T.t[k];
  const size_t base = (size_t)(blockIdx.x - T.first_block[k]) * kAdamPerBlock;
  const float step_size = T.hyper ? T.hyper[1 + t.lr_index] * T.hyper[0] : t.lr * T.c_step;
  const bool vec = ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0);
  auto upd = [&](float& p, float g, float& m, float& v) {
    m = T.b0 * m + (1.0f - T.b0) * g;
    v = T.b1 * v + (1.0f - T.b1) * g * g;
    p = p - m * step_size / (sqrtf(v) + T.eps);
  };
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const size_t i = base + ((size_t)it * 256 + threadIdx.x) * 4;
    if (i >= t.n) break;
    if (vec && i + 4 <= t.n) {
      float4 p = *reinterpret_cast<float4*>(t.p + i), m = *reinterpret_cast<float4*>(t.m + i),
             v = *reinterpret_cast<float4*>(t.v + i);
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
    } else {
      for (size_t j = i; j < i + 4 && j < t.n; ++j) upd(t.p[j], t.g[j], t.m[j], t.v[j]);
    }
  }
}

static int stream_grid(size_t total) {
  size_t b = (total + 255) / 256;
  const size_t cap = 148 * 8;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace tvm

using namespace tvm;

extern "C" int tvm_tv_loss(const float* plane_nchw, int C, int H, int W, float weight, const float* weight_dev,
                           float* loss_accum, float* grad_nchw, void* stream) {
  TVM_REQUIRE(plane_nchw && C > 0 && H > 0 && W > 0, "bad arguments");
  // TVLoss (batch 1): weight * 2 * (h_tv / count_h + w_tv / count_w), count_h = C (H-1) W, count_w = C H (W-1)
  const double ch = (double)C * (H - 1) * W, cw = (double)C * H * (W - 1);
  const float sh = ch > 0 ? (float)(2.0 * weight / ch) : 0.0f, sw = cw > 0 ? (float)(2.0 * weight / cw) : 0.0f;
  k_tv_loss<<<stream_grid((size_t)C * H * W), 256, 0, (cudaStream_t)stream>>>(plane_nchw, C, H, W, sh, sw, weight_dev, loss_accum, grad_nchw);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_tv_loss_batch(const TvmTvJob* jobs_host, int n_jobs, float* loss_accum, void* stream) {
  TVM_REQUIRE(jobs_host && n_jobs > 0 && n_jobs <= TVM_TV_MAX, "tvm_tv_loss_batch takes 1..%d planes", TVM_TV_MAX);
  TvBatch B;
  size_t largest = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const TvmTvJob& j = jobs_host[i];
    TVM_REQUIRE(j.plane_nchw && j.C > 0 && j.H > 0 && j.W > 0, "bad TV job %d", i);
    B.job[i] = j;
    const double ch = (double)j.C * (j.H - 1) * j.W, cw = (double)j.C * j.H * (j.W - 1);
    B.scale_h[i] = ch > 0 ? (float)(2.0 * j.weight / ch) : 0.0f;
    B.scale_w[i] = cw > 0 ? (float)(2.0 * j.weight / cw) : 0.0f;
    largest = std::max(largest, (size_t)j.C * j.H * j.W);
  }
  const int gx = std::max(1, stream_grid(largest) / n_jobs);
  k_tv_loss_batch<<<dim3(gx, n_jobs), 256, 0, (cudaStream_t)stream>>>(B, loss_accum);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_l1_loss(const float* x, size_t n, float weight, const float* weight_dev, float* loss_accum, float* grad,
                           void* stream) {
  TVM_REQUIRE(x && n > 0, "bad arguments");
  k_l1_loss<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(x, n, (float)(weight / (double)n), weight_dev, loss_accum, grad);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_vector_diffs(const float* line_cl, int C, int L, float weight, const float* weight_dev, float* loss_accum,
                                float* grad, void* stream) {
  TVM_REQUIRE(line_cl && C > 1 && C <= 64 && L > 0, "vector_diffs supports 2..64 components");
  k_vector_diffs<<<1, 256, 0, (cudaStream_t)stream>>>(line_cl, C, L, (float)(weight / ((double)C * (C - 1))), weight_dev, loss_accum, grad);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tvm_adam_step(const TvmAdamTensor* tensors_host, int n_tensors, float beta0, float beta1, float eps, int step,
                             const float* hyper_dev, void* stream) {
  TVM_REQUIRE(tensors_host && n_tensors > 0 && step >= 1, "bad arguments");
  for (int s0 = 0; s0 < n_tensors; s0 += TVM_ADAM_MAX_TENSORS) {
    AdamTable T;
    T.n = n_tensors - s0 < TVM_ADAM_MAX_TENSORS ? n_tensors - s0 : TVM_ADAM_MAX_TENSORS;
    unsigned blocks = 0;
    for (int k = 0; k < T.n; ++k) {
      T.t[k] = tensors_host[s0 + k];
      TVM_REQUIRE(T.t[k].p && T.t[k].g && T.t[k].m && T.t[k].v && T.t[k].n > 0, "null / empty Adam tensor");
      T.first_block[k] = blocks;
      blocks += (unsigned)((T.t[k].n + kAdamPerBlock - 1) / kAdamPerBlock);
    }
    T.first_block[T.n] = blocks;
    T.b0 = beta0; T.b1 = beta1; T.eps = eps;
    T.hyper = hyper_dev;
    T.c_step = (float)(sqrt(1.0 - pow((double)beta1, step)) / (1.0 - pow((double)beta0, step)));
    k_adam_multi<<<blocks, 256, 0, (cudaStream_t)stream>>>(T);
    TVM_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
