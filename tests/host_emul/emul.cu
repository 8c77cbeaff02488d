// TEST INFRASTRUCTURE: compiles the kernels' __host__ __device__ per-sample arithmetic
// (jittor-myc-nerfs_b200/csrc/tvm_math.cuh) for the HOST and walks rays sequentially, so that the
// index/layout conventions of the CUDA kernels can be checked against the oracle inside the
// CPU-only build container.  It is never linked into libtvmrender.so and is not a product path.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../jittor-myc-nerfs_b200/csrc/tvm_math.cuh"

using namespace tvm;

// All TvmModel pointers are HOST pointers here (same packed layouts as on the device).
extern "C" int emul_forward(const TvmModel* mp, const float* rays, int n, int S, const float* jitter,
                            uint32_t flags, uint8_t* bbox, uint8_t* valid, uint8_t* app, float* sigma_o,
                            float* weight_o, float* rgb_o, float* rgb_map, float* depth_map) {
  const TvmModel& m = *mp;
  const int Cd = m.n_density, Ca = m.n_app, K = 3 * Ca, F = m.feature_c;
  const int in_c = 2 * m.view_pe * 3 + 2 * m.fea_pe * m.app_dim + 3 + m.app_dim;
  std::vector<float> h(K), x(in_c), y1(F), y2(F);
  for (int ray = 0; ray < n; ++ray) {
    RayMarch r;
    ray_setup(m, rays + 6 * (size_t)ray, jitter, ray, S, r);
    float T = 1.0f, acc = 0.0f, dep = 0.0f, c0 = 0, c1 = 0, c2 = 0;
    for (int k = 0; k < S; ++k) {
      const size_t idx = (size_t)ray * S + k;
      float z = sample_z(m, r, k), p[3];
      bool inside = sample_point(m, r, z, p);
      bool ok = inside;
      if (ok && m.alpha_bits) ok = alpha_mask_test(m, m.alpha_bits, p);
      bbox[idx] = inside;
      valid[idx] = ok;
      float sigma = 0.0f;
      float u[3];
      Axis ax[3];
      if (ok) {
        grid_coords(m, p, u);
        for (int i = 0; i < 3; ++i) ax[i] = axis_taps(u[i], m.grid[i]);
        // density: the adjacent-pair form k_march uses (axis_pair / vm_pair); appearance below: the 4-address form
        AxisPair ap[3];
        for (int i = 0; i < 3; ++i) ap[i] = axis_pair(u[i], m.grid[i]);
        float f = 0.0f;
        for (int kk = 0; kk < 3; ++kk) {
          VmPair t = vm_pair(m, ap, kk, Cd);
          for (int c = 0; c < Cd; c += 4) {
            float4 pv, lv;
            vm_pair_sample4(m.density_plane[kk], m.density_line[kk], t, Cd, c, pv, lv);
            f += pv.x * lv.x + pv.y * lv.y + pv.z * lv.z + pv.w * lv.w;
          }
        }
        sigma = feature2density(m, f);
      }
      float z1 = sample_z(m, r, k + 1);
      float dist = (k < S - 1) ? TVM_MUL(TVM_SUB(z1, z), m.distance_scale) : 0.0f;
      float alpha = TVM_SUB(1.0f, expf(TVM_MUL(-sigma, dist)));
      float w = alpha * T;
      T = T * TVM_ADD(TVM_SUB(1.0f, alpha), 1e-10f);
      acc += w;
      dep += w * z;
      bool a = w > m.weight_thres;
      app[idx] = a;
      sigma_o[idx] = sigma;
      weight_o[idx] = w;
      float rgb[3] = {0, 0, 0};
      if (a) {
        for (int kk = 0; kk < 3; ++kk) {
          VmTaps t = vm_taps(m, ax, kk);
          for (int c = 0; c < Ca; c += 4) {
            float4 pv, lv;
            vm_sample4(m.app_plane[kk], m.app_line[kk], t, Ca, c, pv, lv);
            float* o = &h[kk * Ca + c];
            o[0] = pv.x * lv.x; o[1] = pv.y * lv.y; o[2] = pv.z * lv.z; o[3] = pv.w * lv.w;
          }
        }
        const int pe_f = m.app_dim + 3, pe_v = pe_f + 2 * m.fea_pe * m.app_dim;
        for (int o = 0; o < m.app_dim; ++o) {
          float s = 0.0f;
          for (int j = 0; j < K; ++j) s = fmaf(h[j], m.basis_t[j * 32 + o], s);
          x[o] = s;
          float fr = 1.0f;
          for (int q = 0; q < m.fea_pe; ++q, fr *= 2.0f) {
            x[pe_f + o * m.fea_pe + q] = sinf(s * fr);
            x[pe_f + m.fea_pe * m.app_dim + o * m.fea_pe + q] = cosf(s * fr);
          }
        }
        for (int c = 0; c < 3; ++c) {
          float d = r.d[c];
          x[m.app_dim + c] = d;
          float fr = 1.0f;
          for (int q = 0; q < m.view_pe; ++q, fr *= 2.0f) {
            x[pe_v + c * m.view_pe + q] = sinf(d * fr);
            x[pe_v + 3 * m.view_pe + c * m.view_pe + q] = cosf(d * fr);
          }
        }
        for (int o = 0; o < F; ++o) {
          float s = m.b1[o];
          for (int j = 0; j < in_c; ++j) s = fmaf(x[j], m.w1_t[(size_t)j * F + o], s);
          y1[o] = fmaxf(s, 0.0f);
        }
        for (int o = 0; o < F; ++o) {
          float s = m.b2[o];
          for (int j = 0; j < F; ++j) s = fmaf(y1[j], m.w2_t[(size_t)j * F + o], s);
          y2[o] = fmaxf(s, 0.0f);
        }
        for (int o = 0; o < 3; ++o) {
          float s = m.b3[o];
          for (int j = 0; j < F; ++j) s = fmaf(y2[j], m.w3[(size_t)o * F + j], s);
          rgb[o] = 1.0f / (1.0f + expf(-s));
        }
        c0 = fmaf(w, rgb[0], c0); c1 = fmaf(w, rgb[1], c1); c2 = fmaf(w, rgb[2], c2);
      }
      rgb_o[idx * 3] = rgb[0]; rgb_o[idx * 3 + 1] = rgb[1]; rgb_o[idx * 3 + 2] = rgb[2];
    }
    float bg = (flags & TVM_WHITE_BG) ? 1.0f - acc : 0.0f;
    rgb_map[ray * 3 + 0] = fminf(fmaxf(c0 + bg, 0.0f), 1.0f);
    rgb_map[ray * 3 + 1] = fminf(fmaxf(c1 + bg, 0.0f), 1.0f);
    rgb_map[ray * 3 + 2] = fminf(fmaxf(c2 + bg, 0.0f), 1.0f);
    depth_map[ray] = dep + (1.0f - acc) * r.d[2];
  }
  return 0;
}

// visit[ray][b] = block_maybe(...) for every 32-sample block (the kernels' coarse empty-space pass)
extern "C" int emul_block_maybe(const TvmModel* mp, const float* rays, int n, int S, const float* jitter, uint8_t* visit) {
  const TvmModel& m = *mp;
  const int NB = (S + 31) / 32;
  for (int ray = 0; ray < n; ++ray) {
    RayMarch r;
    ray_setup(m, rays + 6 * (size_t)ray, jitter, ray, S, r);
    for (int b = 0; b < NB; ++b) visit[(size_t)ray * NB + b] = block_maybe(m, r, b, S);
  }
  return 0;
}

// axis_pair vs axis_taps: p0 T[b] + p1 T[b+1] must equal w0 T[i0] + w1 T[i1] bit for bit on any table T
extern "C" int emul_axis_pair_check(const float* u, int n, int size, const float* T, float* out_taps, float* out_pair) {
  for (int i = 0; i < n; ++i) {
    const Axis a = axis_taps(u[i], size);
    const AxisPair p = axis_pair(u[i], size);
    out_taps[i] = TVM_ADD(TVM_MUL(a.w0, T[a.i0]), TVM_MUL(a.w1, T[a.i1]));
    out_pair[i] = TVM_ADD(TVM_MUL(p.p0, T[p.b]), TVM_MUL(p.p1, T[p.b + 1]));
  }
  return 0;
}
