// Backward pass of the TensoRF-VM ray renderer: d rgb_map -> gradients of every parameter
// (row a12 of SURVEY.md §8a; the reference gets this from Jittor autograd, train.py:228,260).
//
// Recompute-based: nothing per-sample is saved by the forward pass except the appearance entry list
// (ray, sample, weight, rgb).  Coordinates are detached (tensoRF.py:212-214,231-233) and depth_map is
// built under no_grad (tensorBase.py:529-531), so only plane/line/basis/MLP parameters get gradients.
//
//   k_bwd_prep   per ray: clamp mask of d rgb_map, S = sum_k dL/dw_k * w_k (closed form from the forward sums)
//   k_app_bwd    per tile of 64 entries: recompute the fp32 head, back-propagate through
//                sigmoid / MLP / positional encoding / basis_mat, accumulate weight gradients
//                (tile-level products, then red.global.add.v4.f32), scatter into app planes/lines
//   k_march_bwd  per ray (one warp, same march as k_march): dL/dw -> dL/dalpha through the transmittance
//                product with ONE forward sweep (suffix sums = total - prefix), -> dL/dsigma -> softplus'
//                -> scatter into density planes/lines with vector atomics
#include "tvm_bwd_simt.cuh"

namespace tvm {

// ------------------------------------------------------------------------------------------------
__global__ void k_bwd_prep(const BwdParams B) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= B.f.n) return;
  const float acc = B.f.ws.acc[ray];
  const float bg = (B.f.flags & TVM_WHITE_BG) ? 1.0f - acc : 0.0f;
  float g[3], tot = 0.0f, gs = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = B.f.ws.rgb_sum[(size_t)ray * 3 + c];
    const float raw = s + bg;
    g[c] = (raw >= 0.0f && raw <= 1.0f) ? B.d_rgb_map[(size_t)ray * 3 + c] : 0.0f;   // clamp(0,1) (tensorBase.py:527)
    tot = fmaf(g[c], s, tot);
    gs += g[c];
  }
  if (!(B.f.flags & TVM_WHITE_BG)) gs = 0.0f;
  // sum_k dL/dw_k * w_k with dL/dw_k = g . rgb_k - gs   (rgb_map = sum w rgb + 1 - sum w)
  // REFTensoRF: + dL/dpenalty * pen_k, penalty = sum w_k pen_k (REFTensoRF.py:236-238)
  if (B.d_penalty) tot = fmaf(*B.d_penalty, B.f.ws.pen_sum[ray], tot);
  float4 o = make_float4(g[0], g[1], g[2], tot - gs * acc);
  reinterpret_cast<float4*>(B.f.ws.bwd_scratch)[ray] = o;
  if (B.f.m.sampling == TVM_SAMPLING_NPP) {
    // rgb_map = clamp(fg) + bg_lambda * bg_rgb_map (nerfplusplus.py:314-317): dL/d bg_lambda = d_rgb_map . bg_rgb_map,
    // unclamped; the gate (:313) passes bg_lambda through or zeroes it.  Stored pre-multiplied by bg_lambda, because
    // d bg_lambda / d alpha_k = -bg_lambda / (1 - alpha_k + 1e-6).
    const float lam = B.f.ws.bg_lambda[ray];
    float gl = 0.0f;
    if (lam > 0.0f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gl = fmaf(B.d_rgb_map[(size_t)ray * 3 + c], B.f.ws.bg_rgb[(size_t)ray * 3 + c], gl);
      gl *= lam;
    }
    B.f.ws.bwd_lam[ray] = gl;
  }
}

// ------------------------------------------------------------------------------------------------
template <int NH>
__global__ void __launch_bounds__(kAppThreads) k_app_bwd(const BwdParams B) {
  constexpr bool REF = NH == TVM_REF_HEAD_LD;
  extern __shared__ __align__(16) float smem[];
  const FwdParams& P = B.f;
  const TvmModel& m = P.m;
  const int st = P.st;
  float* H = smem;                        // appearance vector h            -> later d h
  float* X = smem + kAppTile * st;        // MLP input x
  float* Y1 = smem + 2 * kAppTile * st;   // layer-1 output                 -> dz1 -> d feat
  float* Y2 = smem + 3 * kAppTile * st;   // layer-2 output                 -> dz2 -> d x
  float* D3 = smem + 4 * kAppTile * st;   // [64][4] dz3
  float* HD = D3 + kAppTile * 4;          // [64][8]  REF: {rgb_d[3], tint, raw normal[3], d.n} from the forward recompute
  float* HG = HD + kAppTile * 8;          // [64][8]  REF: {w g [3], sigmoid [3], ...}
  const int tid = threadIdx.x;
  const int row = tid & (kAppTile - 1), part = tid >> 6;
  const int Ca = m.n_app, K0 = 3 * Ca;
  const uint32_t n_ent = *P.ws.n_entries;
  const uint32_t n_tiles = (n_ent + kAppTile - 1) / kAppTile;
  const float4* gray = reinterpret_cast<const float4*>(P.ws.bwd_scratch);

  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t tile_base = tile * kAppTile;
    const uint32_t e = tile_base + row;
    // ---- forward recompute ---------------------------------------------------------------------
    app_gather_tile(P, tile_base, n_ent, H, X, st);
    __syncthreads();
    app_basis_pe<NH>(P, H, X, HD, st, tile_base, n_ent);
    __syncthreads();
    app_dense<true>(m.w1_t, m.b1, X, P.in_mlp_c, Y1, st);
    __syncthreads();
    app_dense<true>(m.w2_t, m.b2, Y1, kFeatureC, Y2, st);
    __syncthreads();
    if (part < 3) {
      float dz = 0.0f;
      if (e < n_ent) {
        const float s = 1.0f / (1.0f + expf(-app_out_logit(m, Y2, st, row, part)));
        const float4 g = gray[P.ws.ent[e].x];
        const float gc = part == 0 ? g.x : (part == 1 ? g.y : g.z);
        dz = P.ws.ent_w[e] * gc * s * (1.0f - s);     // d/d logit of w * rgb . g
        if (REF) {
          // rgb = tint * rgb_s + rgb_d (REFTensoRF.py:232; rgb_s = sigmoid > 0, so the clamp is the identity)
          dz *= HD[row * 8 + 3];
          HG[row * 8 + part] = P.ws.ent_w[e] * gc;
          HG[row * 8 + 3 + part] = s;
        }
      } else if (REF) {
        HG[row * 8 + part] = 0.0f;
        HG[row * 8 + 3 + part] = 0.0f;
      }
      D3[row * 4 + part] = dz;
    }
    __syncthreads();
    // ---- layer 3: dW3, db3, dY2 -> dz2 (in place of Y2) ----------------------------------------------
    {
      const int j = tid & (kFeatureC - 1);
      const int o_lo = (tid >> 7) ? 2 : 0, o_hi = (tid >> 7) ? 3 : 2;
      float a0 = 0.0f, a1 = 0.0f;
      for (int r = 0; r < kAppTile; ++r) {
        const float y = Y2[r * st + j];
        a0 = fmaf(D3[r * 4 + o_lo], y, a0);
        if (o_hi - o_lo == 2) a1 = fmaf(D3[r * 4 + o_lo + 1], y, a1);
      }
      atomicAdd(B.g.w3 + o_lo * kFeatureC + j, a0);
      if (o_hi - o_lo == 2) atomicAdd(B.g.w3 + (o_lo + 1) * kFeatureC + j, a1);
      if (tid < 3) {
        float a = 0.0f;
        for (int r = 0; r < kAppTile; ++r) a += D3[r * 4 + tid];
        atomicAdd(B.g.b3 + tid, a);
      }
    }
    __syncthreads();
    {
      const float d0 = D3[row * 4 + 0], d1 = D3[row * 4 + 1], d2 = D3[row * 4 + 2];
      float* y = Y2 + row * st + part * 32;
      const float* w = m.w3 + part * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 yv = lds4(y + i);
        const float4 w0 = ldg4(w + i), w1 = ldg4(w + kFeatureC + i), w2 = ldg4(w + 2 * kFeatureC + i);
        float4 o;
        o.x = yv.x > 0.0f ? fmaf(d0, w0.x, fmaf(d1, w1.x, d2 * w2.x)) : 0.0f;
        o.y = yv.y > 0.0f ? fmaf(d0, w0.y, fmaf(d1, w1.y, d2 * w2.y)) : 0.0f;
        o.z = yv.z > 0.0f ? fmaf(d0, w0.z, fmaf(d1, w1.z, d2 * w2.z)) : 0.0f;
        o.w = yv.w > 0.0f ? fmaf(d0, w0.w, fmaf(d1, w1.w, d2 * w2.w)) : 0.0f;
        *reinterpret_cast<float4*>(y + i) = o;
      }
    }
    __syncthreads();
    // ---- layer 2: dW2, db2, dY1 -> dz1 (in place of Y1) ----------------------------------------------
    wgrad_tile_128(Y1, Y2, st, kFeatureC, B.g.w2_t, kFeatureC);
    bgrad_tile(Y2, st, B.g.b2);
    __syncthreads();
    app_dense_bwd<true>(m.w2_t, Y2, kFeatureC, Y1, Y1, st);
    __syncthreads();
    // ---- layer 1: dW1, db1, d x (into Y2) ----------------------------------------------------------
    wgrad_tile_128(X, Y1, st, P.in_mlp_c, B.g.w1_t, kFeatureC);
    bgrad_tile(Y1, st, B.g.b1);
    app_dense_bwd<false>(m.w1_t, Y1, P.in_mlp_c, Y2, nullptr, st);
    __syncthreads();
    // ---- positional encoding: d head outputs (into Y1[.., 0..NH)) -------------------------------------------
    {
      const float* x = X + row * st;
      const float* dx = Y2 + row * st;
      const int c_feat = col_feat(m), c_dir = col_dir(m);
      const int pe_f = c_dir + 3;
      const int n_f = m.fea_pe * m.app_dim;
      constexpr int NP = NH / 4;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int o = part * NP + i;
        float df = 0.0f;
        if (o < m.app_dim) {
          df = dx[c_feat + o];
          float fr = 1.0f;
          for (int q = 0; q < m.fea_pe; ++q, fr *= 2.0f) {
            const int si = pe_f + o * m.fea_pe + q, ci = si + n_f;
            df += fr * (dx[si] * x[ci] - dx[ci] * x[si]);     // d sin = cos, d cos = -sin
          }
        }
        Y1[row * st + o] = df;
      }
      if (REF) __syncthreads();
      if (REF && part == 3) {
        // REFTensoRF.py:216-238 backwards: reflection, dot product, normalisation, tint / diffuse mix, penalty
        const uint32_t e = tile_base + row;
        float dv[3] = {0.0f, 0.0f, 0.0f}, dspec = 0.0f, drd[3] = {0.0f, 0.0f, 0.0f};
        if (e < n_ent) {
          const int pe_v = pe_f + 2 * n_f, n_v = 3 * m.view_pe;
          float dr[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float a = dx[c_dir + c];
            float fr = 1.0f;
            for (int q = 0; q < m.view_pe; ++q, fr *= 2.0f) {
              const int si = pe_v + c * m.view_pe + q, ci = si + n_v;
              a += fr * (dx[si] * x[ci] - dx[ci] * x[si]);
            }
            dr[c] = a;                                        // d reflection
          }
          const float* hd = HD + row * 8;
          const float* hg = HG + row * 8;
          const float v0 = hd[4], v1 = hd[5], v2 = hd[6], dot = hd[7];
          const float n2 = v0 * v0 + v1 * v1 + v2 * v2;
          const float inv = 1.0f / sqrtf(fmaxf(n2, 1e-30f));
          const float nh[3] = {v0 * inv, v1 * inv, v2 * inv};
          const uint32_t ray = P.ws.ent[e].x;
          const float d[3] = {-P.rays[6 * (size_t)ray + 3], -P.rays[6 * (size_t)ray + 4], -P.rays[6 * (size_t)ray + 5]};
          // x[0] = -dot; reflection = 2 dot n - d; penalty = sum w relu(-dot)^2
          float ddot = -dx[0] + 2.0f * (nh[0] * dr[0] + nh[1] * dr[1] + nh[2] * dr[2]);
          if (B.d_penalty) ddot -= *B.d_penalty * P.ws.ent_w[e] * 2.0f * fmaxf(-dot, 0.0f);
          float dn[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) dn[c] = 2.0f * dot * dr[c] + ddot * d[c];
          const float proj = nh[0] * dn[0] + nh[1] * dn[1] + nh[2] * dn[2];
          if (n2 > 1e-30f) {
#pragma unroll
            for (int c = 0; c < 3; ++c) dv[c] = (dn[c] - nh[c] * proj) * inv;
          }
          const float dtint = hg[0] * hg[3] + hg[1] * hg[4] + hg[2] * hg[5];
          dspec = hd[3] > 0.0f ? dtint : 0.0f;                // tint = relu(specular_linear(h))
          drd[0] = hg[0]; drd[1] = hg[1]; drd[2] = hg[2];     // d rgb_d = w g
        }
        float* y = Y1 + row * st + m.app_dim;                 // head order: normal 3 | diffuse 3 | specular | rho
        y[0] = dv[0]; y[1] = dv[1]; y[2] = dv[2];
        y[3] = drd[0]; y[4] = drd[1]; y[5] = drd[2];
        y[6] = dspec;
        y[7] = 0.0f;                                          // rho only feeds the unused 1/rho argument (REFTensoRF.py:226)
        for (int o = m.app_dim + 8; o < NH; ++o) Y1[row * st + o] = 0.0f;
      }
    }
    __syncthreads();
    // ---- basis_mat (+ heads): d basis, d head bias, d h (into X) ----------------------------------------------
    wgrad_tile_heads<NH>(H, Y1, st, K0, B.g.basis_t);
    if (REF && tid < 8) {
      float a = 0.0f;
      for (int r = 0; r < kAppTile; ++r) a += Y1[r * st + m.app_dim + tid];
      atomicAdd(B.g.head_bias + m.app_dim + tid, a);
    }
    {
      float df[NH];
#pragma unroll
      for (int i = 0; i < NH; i += 4) {
        const float4 v = lds4(Y1 + row * st + i);
        df[i] = v.x; df[i + 1] = v.y; df[i + 2] = v.z; df[i + 3] = v.w;
      }
      const int JP = (K0 / 4 + 3) & ~3;
      const int jbeg = part * JP, jend = min(K0, jbeg + JP);
      for (int j = jbeg; j < jend; ++j) {
        const float* bt = m.basis_t + j * NH;
        float a = 0.0f;
#pragma unroll
        for (int i = 0; i < NH; i += 4) {
          const float4 w = ldg4(bt + i);
          a = fmaf(df[i], w.x, fmaf(df[i + 1], w.y, fmaf(df[i + 2], w.z, fmaf(df[i + 3], w.w, a))));
        }
        X[row * st + j] = a;
      }
    }
    __syncthreads();
    // ---- scatter d h into the appearance planes / lines (4 lanes per entry, as in the gather) ----------
    {
      const int warp = tid >> 5, lane = tid & 31;
      const int grow = warp * 8 + (lane >> 2), q = lane & 3;
      const uint32_t ge = tile_base + grow;
      if (ge < n_ent) {
        const uint2 en = P.ws.ent[ge];
        float u[3], dir[3];
        entry_coords(m, P.rays, P.jitter, en.x, en.y, P.S, u, dir);
        Axis ax[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) ax[i] = axis_taps(u[i], m.grid[i]);
        const float* dh = X + grow * st;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const VmTaps t = vm_taps(m, ax, kk);
          float* gp = B.g.app_plane[kk];
          float* gl = B.g.app_line[kk];
          for (int c = q * 4; c < Ca; c += 16) {
            float4 pv, lv;
            vm_sample4(m.app_plane[kk], m.app_line[kk], t, Ca, c, pv, lv);
            const float4 d = lds4(dh + kk * Ca + c);
            const float px = d.x * lv.x, py = d.y * lv.y, pz = d.z * lv.z, pw = d.w * lv.w;   // d plane value
            const float lx = d.x * pv.x, ly = d.y * pv.y, lz = d.z * pv.z, lw = d.w * pv.w;   // d line value
            red_add_v4(gp + (size_t)t.o00 * Ca + c, px * t.nw, py * t.nw, pz * t.nw, pw * t.nw);
            red_add_v4(gp + (size_t)t.o01 * Ca + c, px * t.ne, py * t.ne, pz * t.ne, pw * t.ne);
            red_add_v4(gp + (size_t)t.o10 * Ca + c, px * t.sw, py * t.sw, pz * t.sw, pw * t.sw);
            red_add_v4(gp + (size_t)t.o11 * Ca + c, px * t.se, py * t.se, pz * t.se, pw * t.se);
            red_add_v4(gl + (size_t)t.l0 * Ca + c, lx * t.lw0, ly * t.lw0, lz * t.lw0, lw * t.lw0);
            red_add_v4(gl + (size_t)t.l1 * Ca + c, lx * t.lw1, ly * t.lw1, lz * t.lw1, lw * t.lw1);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// k_march_bwd: same march / control flow as k_march (tvm_forward.cu), one warp per ray
// ------------------------------------------------------------------------------------------------
constexpr int kMarchWarps = 8;

__global__ void __launch_bounds__(kMarchWarps * 32) k_march_bwd(const BwdParams B) {
  __shared__ float s_u[kMarchWarps][32][3];
  __shared__ float s_f[kMarchWarps][32];
  const FwdParams& P = B.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kMarchWarps + warp;
  if (ray >= P.n) return;
  const TvmModel& m = P.m;
  const uint32_t lt_mask = (1u << lane) - 1u;

  RayMarch r;
  ray_setup(m, P.rays + 6 * (size_t)ray, P.jitter, ray, P.S, r);
  const float4 g = reinterpret_cast<const float4*>(P.ws.bwd_scratch)[ray];
  const float gs = (P.flags & TVM_WHITE_BG) ? g.x + g.y + g.z : 0.0f;
  const float total = g.w;
  const float dpen = B.d_penalty ? *B.d_penalty : 0.0f;
  const float glam = m.sampling == TVM_SAMPLING_NPP ? P.ws.bwd_lam[ray] : 0.0f;   // NeRF++: bg_lambda * dL/d bg_lambda
  if (g.x == 0.0f && g.y == 0.0f && g.z == 0.0f && dpen == 0.0f && glam == 0.0f) return;   // nothing flows into this ray

  const int C = m.n_density, S = P.S;
  const bool ert = !(P.flags & TVM_NO_ERT);
  const int q = lane & 3;
  float T = 1.0f, carry = 0.0f;
  bool seen = false;

  // coarse pass (one lane per 32-sample block): blocks that cannot hold a valid sample are never visited
  const bool use_visit = P.NB <= 64;
  unsigned long long visit = 0ull;
  if (use_visit) {
    for (int b0 = 0; b0 < P.NB; b0 += 32) {
      const int bb = b0 + lane;
      const bool maybe = bb < P.NB && block_maybe(m, r, bb, S);
      visit |= (unsigned long long)__ballot_sync(0xffffffffu, maybe) << b0;
    }
  }
  int b = -1;
  while (true) {
    if (use_visit) {
      if (!visit) break;
      b = __ffsll((long long)visit) - 1;
      visit &= visit - 1;
    } else if (++b >= P.NB) {
      break;
    }
    const int k = b * 32 + lane;
    const float z = sample_z(m, r, k);
    float p[3];
    bool inside = sample_point(m, r, z, p) && (k < S);
    const uint32_t in_bits = __ballot_sync(0xffffffffu, inside);
    if (in_bits == 0) {
      if (seen) break;
      continue;
    }
    seen = true;
    bool valid = inside;
    if (m.alpha_bits != nullptr && inside) valid = alpha_mask_test(m, m.alpha_bits, p);
    const uint32_t v_bits = __ballot_sync(0xffffffffu, valid);
    const int nv = __popc(v_bits);
    const int rank = __popc(v_bits & lt_mask);

    float f = 0.0f, sigma = 0.0f;
    if (v_bits != 0) {
      if (valid) {
        float u[3];
        grid_coords(m, p, u);
        s_u[warp][rank][0] = u[0];
        s_u[warp][rank][1] = u[1];
        s_u[warp][rank][2] = u[2];
      }
      __syncwarp();
      for (int gi = 0; gi < nv; gi += 8) {
        const int j = gi + (lane >> 2);
        float part = 0.0f;
        if (j < nv) {
          Axis ax[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) ax[i] = axis_taps(s_u[warp][j][i], m.grid[i]);
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const VmTaps t = vm_taps(m, ax, kk);
            for (int c = q * 4; c < C; c += 16) {
              float4 pv, lv;
              vm_sample4(m.density_plane[kk], m.density_line[kk], t, C, c, pv, lv);
              part += pv.x * lv.x + pv.y * lv.y + pv.z * lv.z + pv.w * lv.w;
            }
          }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (j < nv && q == 0) s_f[warp][j] = part;
      }
      __syncwarp();
      if (valid) {
        f = s_f[warp][rank];
        sigma = feature2density(m, f);
      }
      __syncwarp();
    }

    const float z1 = sample_z(m, r, k + 1);
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
    const float Tk = T * excl;
    const float w = alpha * Tk;
    T = T * __shfl_sync(0xffffffffu, pref, 31);

    // dL/dw_k = g . rgb_k (weighted samples only) - gs
    const uint32_t a_bits = P.ws.blk_mask[(size_t)ray * P.NB + b];
    float dw = -gs;
    if ((a_bits >> lane) & 1u) {
      const uint32_t e = P.ws.blk_base[(size_t)ray * P.NB + b] + __popc(a_bits & lt_mask);
      const float* c3 = P.ws.ent_rgb + (size_t)e * 3;
      dw += g.x * c3[0] + g.y * c3[1] + g.z * c3[2];
      if (dpen != 0.0f) dw = fmaf(dpen, P.ws.ent_pen[e], dw);
    }
    // inclusive prefix of dw*w; the suffix sum_{j>k} dw_j w_j is total - prefix
    float ps = dw * w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, ps, o);
      if (lane >= o) ps += t;
    }
    const float suffix = total - (carry + ps);
    carry += __shfl_sync(0xffffffffu, ps, 31);
    float dalpha = dw * Tk - suffix / v;
    if (glam != 0.0f && k < S) dalpha -= glam / TVM_ADD(TVM_SUB(1.0f, alpha), 1e-6f);   // through prod(1 - alpha + 1e-6)
    const float dsigma = dalpha * dist * (1.0f - alpha);
    const float dfeat = valid ? dsigma * feature2density_grad(m, f) : 0.0f;

    if (v_bits != 0) {
      if (valid) s_f[warp][rank] = dfeat;
      __syncwarp();
      for (int gi = 0; gi < nv; gi += 8) {
        const int j = gi + (lane >> 2);
        if (j < nv) {
          const float d = s_f[warp][j];
          if (d != 0.0f) {
            Axis ax[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) ax[i] = axis_taps(s_u[warp][j][i], m.grid[i]);
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
              const VmTaps t = vm_taps(m, ax, kk);
              float* gp = B.g.density_plane[kk];
              float* gl = B.g.density_line[kk];
              for (int c = q * 4; c < C; c += 16) {
                float4 pv, lv;
                vm_sample4(m.density_plane[kk], m.density_line[kk], t, C, c, pv, lv);
                const float px = d * lv.x, py = d * lv.y, pz = d * lv.z, pw = d * lv.w;
                const float lx = d * pv.x, ly = d * pv.y, lz = d * pv.z, lw = d * pv.w;
                red_add_v4(gp + (size_t)t.o00 * C + c, px * t.nw, py * t.nw, pz * t.nw, pw * t.nw);
                red_add_v4(gp + (size_t)t.o01 * C + c, px * t.ne, py * t.ne, pz * t.ne, pw * t.ne);
                red_add_v4(gp + (size_t)t.o10 * C + c, px * t.sw, py * t.sw, pz * t.sw, pw * t.sw);
                red_add_v4(gp + (size_t)t.o11 * C + c, px * t.se, py * t.se, pz * t.se, pw * t.se);
                red_add_v4(gl + (size_t)t.l0 * C + c, lx * t.lw0, ly * t.lw0, lz * t.lw0, lw * t.lw0);
                red_add_v4(gl + (size_t)t.l1 * C + c, lx * t.lw1, ly * t.lw1, lz * t.lw1, lw * t.lw1);
              }
            }
          }
        }
      }
      __syncwarp();
    }
    if (ert && T < kErtEps) break;
    if (!(in_bits >> 31)) break;
  }
}

int fill_fwd_params(FwdParams& P, const TvmModel* m, const float* rays, int n, int S, const float* jitter,
                    uint32_t flags, void* ws, size_t ws_bytes, bool bounded_ok);   // tvm_forward.cu
int launch_app_bwd_tc(const BwdParams& B, int num_sms, cudaStream_t stream);   // tvm_bwd_tc.cu

}  // namespace tvm

using namespace tvm;

namespace tvm {
int launch_bg_bwd(const BwdParams& B, const TvmBgGrads& bg_grads, int num_sms, cudaStream_t stream);   // tvm_bg_bwd.cu
}

// shared body of tvm_backward (bg_host == NULL) and tvm_backward_npp
static int backward_impl(const TvmModel* m_host, const TvmBgNet* bg_host, const float* rays, int n_rays, int n_samples,
                         const float* jitter, const float* bg_rand, uint32_t flags, const float* d_rgb_map,
                         const float* d_penalty, const TvmGrads* grads_host, const TvmBgGrads* bg_grads_host, void* ws,
                         size_t ws_bytes, void* stream_, const TvmGradExchange* xchg = nullptr) {
  cudaStream_t stream = (cudaStream_t)stream_;
  BwdParams B;
  if (xchg) {
    TVM_REQUIRE(xchg->comm && xchg->side_stream && xchg->side_stream != stream_, "TvmGradExchange needs a communicator and a second stream");
    TVM_REQUIRE(xchg->split_floats <= xchg->total_floats && (xchg->split_floats & 3) == 0 && (xchg->total_floats & 3) == 0,
                "TvmGradExchange: split and total must be multiples of 4 floats");
  }
  if (bg_host) flags &= ~TVM_WHITE_BG;          // the foreground of NerfPlusPlus renders on black (nerfplusplus.py:274)
  if (int rc = fill_fwd_params(B.f, m_host, rays, n_rays, n_samples, jitter, flags, ws, ws_bytes, false)) return rc;
  TVM_REQUIRE(d_rgb_map && grads_host, "null argument");
  TVM_REQUIRE(!(flags & TVM_EVAL_ONLY), "TVM_EVAL_ONLY is a tvm_forward flag: a forward launched with it leaves no stash for the backward pass");
  TVM_REQUIRE((m_host->sampling == TVM_SAMPLING_NPP) == (bg_host != nullptr),
              "TVM_SAMPLING_NPP models go through tvm_backward_npp, all others through tvm_backward");
  if (bg_host) {
    TVM_REQUIRE(jitter && bg_rand && bg_grads_host, "tvm_backward_npp needs fg_rand, bg_rand and TvmBgGrads");
    B.f.bg = *bg_host;
    B.f.bg_rand = bg_rand;
  }
  const bool ref = m_host->variant == TVM_VARIANT_REF;
  B.d_rgb_map = d_rgb_map;
  B.d_penalty = ref ? d_penalty : nullptr;
  B.g = *grads_host;
  for (int k = 0; k < 3; ++k)
    TVM_REQUIRE(B.g.density_plane[k] && B.g.density_line[k] && B.g.app_plane[k] && B.g.app_line[k], "null gradient pointer");
  TVM_REQUIRE(B.g.basis_t && B.g.w1_t && B.g.b1 && B.g.w2_t && B.g.b2 && B.g.w3 && B.g.b3, "null gradient pointer");
  TVM_REQUIRE(!ref || B.g.head_bias, "TVM_VARIANT_REF needs TvmGrads.head_bias");

  const int phase = xchg ? xchg->phase : 0;       // 0: everything; 1: prep + appearance half; 2: density half
  TVM_REQUIRE(phase >= 0 && phase <= 2, "TvmGradExchange.phase must be 0, 1 or 2");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (phase != 2) {
  k_bwd_prep<<<(n_rays + 255) / 256, 256, 0, stream>>>(B);
  TVM_CHECK_CUDA(cudaGetLastError());
  if ((flags & TVM_MLP_MASK) == TVM_MLP_BF16 || ((flags & TVM_MLP_MASK) == TVM_MLP_FP16 && m_host->tc_weights_bwd)) {
    // appearance backward on the tensor cores (bf16 operands, fp32 accumulation; gradients to ~1e-2 relative)
    ProfileScope prof(TVM_STAGE_BWD_APP, stream);
    if (int rc = launch_app_bwd_tc(B, sms, stream)) return rc;
  } else {
    const size_t smem = ((size_t)kAppTile * 4 * B.f.st + kAppTile * (4 + 8 + 8)) * sizeof(float);
    TVM_REQUIRE(smem <= 220 * 1024, "appearance backward tile does not fit shared memory");
    auto kern = ref ? k_app_bwd<TVM_REF_HEAD_LD> : k_app_bwd<32>;
    TVM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfileScope prof(TVM_STAGE_BWD_APP, stream);
    kern<<<sms, kAppThreads, smem, stream>>>(B);
  }
  TVM_CHECK_CUDA(cudaGetLastError());
  }
  cudaEvent_t ev_join = nullptr;
  if (xchg && phase != 2) {
    // every appearance gradient is final: exchange [split, total) on the side stream WHILE k_march_bwd scatters the density
    // gradients on this one (fork / join through events: inside a stream capture these become graph edges)
    cudaStream_t side = (cudaStream_t)xchg->side_stream;
    cudaEvent_t ev_fork;
    TVM_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    TVM_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    TVM_CHECK_CUDA(cudaEventRecord(ev_fork, stream));
    TVM_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    TVM_CHECK_CUDA(cudaEventDestroy(ev_fork));
    if (int rc = tvm_allreduce_sum(xchg->comm, xchg->split_floats, xchg->total_floats - xchg->split_floats, xchg->n_ctas_overlapped, side)) return rc;
    TVM_CHECK_CUDA(cudaEventRecord(ev_join, side));
    if (phase == 1) {           // the caller joins the side stream itself (and may run the appearance tail of the step first)
      TVM_CHECK_CUDA(cudaEventDestroy(ev_join));
      return 0;
    }
  }
  {
    ProfileScope prof(TVM_STAGE_BWD_MARCH, stream);
    k_march_bwd<<<(n_rays + kMarchWarps - 1) / kMarchWarps, kMarchWarps * 32, 0, stream>>>(B);
  }
  TVM_CHECK_CUDA(cudaGetLastError());
  if (xchg && phase == 0) {
    TVM_CHECK_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
    TVM_CHECK_CUDA(cudaEventDestroy(ev_join));
    if (int rc = tvm_allreduce_sum(xchg->comm, 0, xchg->split_floats, xchg->n_ctas, stream)) return rc;
  }
  if (xchg && phase == 2) {
    // density half on the side stream as well, behind the scatter: the caller's stream stays free for the appearance tail
    cudaStream_t side = (cudaStream_t)xchg->side_stream;
    cudaEvent_t ev;
    TVM_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    TVM_CHECK_CUDA(cudaEventRecord(ev, stream));
    TVM_CHECK_CUDA(cudaStreamWaitEvent(side, ev, 0));
    TVM_CHECK_CUDA(cudaEventDestroy(ev));
    if (int rc = tvm_allreduce_sum(xchg->comm, 0, xchg->split_floats, xchg->n_ctas_overlapped, side)) return rc;
  }
  if (bg_host) {
    ProfileScope prof(TVM_STAGE_BWD_BG, stream);
    if (int rc = launch_bg_bwd(B, *bg_grads_host, sms, stream)) return rc;
  }
  return 0;
}

extern "C" int tvm_backward(const TvmModel* m_host, const float* rays, int n_rays, int n_samples,
                            const float* jitter, uint32_t flags, const float* rgb_map, const float* d_rgb_map,
                            const float* d_penalty, const TvmGrads* grads_host, void* ws, size_t ws_bytes, void* stream_) {
  (void)rgb_map;
  return backward_impl(m_host, nullptr, rays, n_rays, n_samples, jitter, nullptr, flags, d_rgb_map, d_penalty, grads_host,
                       nullptr, ws, ws_bytes, stream_);
}

extern "C" int tvm_backward_dp(const TvmModel* m_host, const float* rays, int n_rays, int n_samples, const float* jitter,
                               uint32_t flags, const float* rgb_map, const float* d_rgb_map, const float* d_penalty,
                               const TvmGrads* grads_host, void* ws, size_t ws_bytes, const TvmGradExchange* xchg, void* stream_) {
  (void)rgb_map;
  TVM_REQUIRE(xchg != nullptr, "null TvmGradExchange");
  return backward_impl(m_host, nullptr, rays, n_rays, n_samples, jitter, nullptr, flags, d_rgb_map, d_penalty, grads_host,
                       nullptr, ws, ws_bytes, stream_, xchg);
}

extern "C" int tvm_backward_npp(const TvmModel* m_host, const TvmBgNet* bg_host, const float* rays, int n_rays,
                                int n_samples, const float* fg_rand, const float* bg_rand, uint32_t flags,
                                const float* rgb_map, const float* d_rgb_map, const TvmGrads* grads_host,
                                const TvmBgGrads* bg_grads_host, void* ws, size_t ws_bytes, void* stream_) {
  (void)rgb_map;
  TVM_REQUIRE(bg_host != nullptr, "null TvmBgNet");
  return backward_impl(m_host, bg_host, rays, n_rays, n_samples, fg_rand, bg_rand, flags, d_rgb_map, nullptr, grads_host,
                       bg_grads_host, ws, ws_bytes, stream_);
}
