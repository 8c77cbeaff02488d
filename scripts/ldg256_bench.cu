// Microbenchmark (diagnostic, not part of the library): the density-tap access shape of k_march with 128-bit loads on the plain
// channels-last layout (texel pair = two 64-byte segments, two LDG.128 per lane) against 256-bit loads on pair records (one
// 128-byte record per texel pair, one LDG.256 per lane).  Random pairs inside a window of `span` texels per CTA (small window =
// L1 hits, whole buffer = L2 hits).  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/ldg256_bench scripts/ldg256_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct f8 { float v[8]; };
__device__ __forceinline__ f8 ldg256(const void* p) {
  f8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}

template <bool WIDE>
__global__ void __launch_bounds__(128, 8) k(const float4* __restrict__ buf, uint32_t n_texels, uint32_t span, int iters, float* sink) {
  const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, q = threadIdx.x & 3;
  uint32_t state = group * 747796405u + 2891336453u;
  const uint32_t base = (uint32_t)(((uint64_t)(blockIdx.x * 2654435761u) * (n_texels - span)) >> 32) & ~1u;
  float acc = 0.0f;
  for (int it = 0; it < iters; ++it) {
    if (WIDE) {
      f8 v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        state = state * 1664525u + 1013904223u;
        const uint32_t texel = (base + (uint32_t)(((uint64_t)(state >> 4) * span) >> 28)) & ~1u;     // record = aligned texel pair: 128 B
        v[t] = ldg256(reinterpret_cast<const char*>(buf) + (size_t)texel * 64 + q * 32);
      }
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[t].v[j];
    } else {
      float4 v[18];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        state = state * 1664525u + 1013904223u;
        const uint32_t texel = base + (uint32_t)(((uint64_t)(state >> 4) * span) >> 28);                // any texel: pair straddles lines half of the time
        v[2 * t] = __ldg(buf + (size_t)texel * 4 + q);
        v[2 * t + 1] = __ldg(buf + (size_t)texel * 4 + 4 + q);
      }
#pragma unroll
      for (int t = 0; t < 18; ++t) acc += v[t].x + v[t].y + v[t].z + v[t].w;
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const size_t bytes = 17u << 20;
  float4* buf; float* sink;
  cudaMalloc(&buf, bytes + 4096); cudaMalloc(&sink, 16);
  cudaMemset(buf, 0, bytes + 4096);
  const uint32_t n_texels = bytes / 64;
  const int blocks = 148 * 8 * 8, iters = 64;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (uint32_t span : {256u, 2048u, n_texels - 2}) {
    for (int wide = 0; wide < 2; ++wide) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        if (wide) k<true><<<blocks, 128>>>(buf, n_texels, span, iters, sink); else k<false><<<blocks, 128>>>(buf, n_texels, span, iters, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      const double pairs = (double)blocks * 32 * iters * 9;     // texel pairs gathered (4 lanes each)
      printf("span %8u texels  %s: %.3f ms  %.1f G pair-taps/s  %.2f TB/s\n", span, wide ? "LDG.256 pair records" : "2 x LDG.128 plain     ", best,
             pairs / best * 1e-6, pairs * 128 / best * 1e-9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
