// Gradient exchange of the data-parallel training step (the reference's train.py:260 is single-process):
// an in-place SUM all-reduce of the flat packed fp32 gradient buffer across the GPUs of one NVSwitch node, written
// directly over peer memory instead of calling NCCL.
//
// Two-shot, one kernel, no host involvement (plain kernel nodes: capturable into the training step's CUDA graph):
//   barrier A   every rank's gradients are complete (the kernels in front of this one in its stream have finished)
//   reduce      rank r owns slice r of the buffer.  NVLS path: ONE multimem.ld_reduce per 16 bytes returns the sum over
//               all peers, added inside the switch, and ONE multimem.st broadcasts it to every peer -- the GPU receives
//               1/world of the buffer for the reduction instead of (world-1)/world.  P2P path (no multicast address):
//               the slice is summed from `world` peer loads and pushed with `world` peer stores.
//   barrier B   every peer has finished writing into this rank's buffer
// A barrier is per CTA: CTA b of rank r exchanges flags with CTA b of every peer (release/acquire at system scope, a
// monotonically increasing epoch as the flag value, so the signal pad is never reset); all CTAs of the grid are
// co-resident (grid <= SM count), and CTA b touches the same byte ranges on every rank, which is all the ordering the
// data needs.  The sum is the same on every rank bit for bit (one rank reduces a slice, all receive that result).
#include "tvm_common.cuh"

namespace tvm {
namespace ar {

constexpr int kThreads = 512;
constexpr int kMaxWorld = 16;

struct Params {
  float* bufs[kMaxWorld];          // peer pointers to the symmetric buffer (index = rank)
  uint32_t* signals[kMaxWorld];    // peer pointers to the signal pads
  float* multicast;                // NVLS address of the buffer, or nullptr
  uint32_t* epoch;                 // device words {calls completed, CTAs finished}: the last CTA of a launch bumps the first
  int rank, world;
  size_t offset4, n4;              // range to reduce, in float4 units
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {      // never from a stale L1 line of an earlier step
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float4* p, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flags of phase ph: signals[rank][(ph * gridDim.x + cta) * world + sender]
__device__ __forceinline__ void peer_barrier(const Params& P, int ph, uint32_t epoch) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < P.world) {
    const int p = threadIdx.x;
    const size_t slot = ((size_t)ph * gridDim.x + blockIdx.x) * P.world;
    st_release_sys(P.signals[p] + slot + P.rank, epoch);
    const uint32_t* mine = P.signals[P.rank] + slot + p;
    uint32_t spin = 0;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spin > (1u << 28)) __trap();      // a peer that never arrives must fail the launch, not hang the node
    }
  }
  __syncthreads();
}

template <bool NVLS>
__global__ void __launch_bounds__(kThreads, 1) k_allreduce(const Params P) {
  // flag value of this call: calls completed so far + 1 (launches of one rank are stream-ordered, so every CTA reads the
  // counter before the last CTA of this launch advances it)
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(P.epoch) + 1u;
  peer_barrier(P, 0, epoch);
  // slice of this rank, split evenly over the CTAs
  const size_t per_rank = (P.n4 + P.world - 1) / P.world;
  const size_t s0 = P.offset4 + per_rank * P.rank;
  const size_t s1 = min(P.offset4 + P.n4, s0 + per_rank);
  const size_t stride = (size_t)gridDim.x * kThreads;
  if (NVLS) {
    float4* mc = reinterpret_cast<float4*>(P.multicast);
    size_t i = s0 + (size_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + 3 * stride < s1; i += 4 * stride) {           // four independent switch round trips in flight per thread
      const float4 a = mc_ld_reduce(mc + i), b = mc_ld_reduce(mc + i + stride), c = mc_ld_reduce(mc + i + 2 * stride),
                   d = mc_ld_reduce(mc + i + 3 * stride);
      mc_st(mc + i, a);
      mc_st(mc + i + stride, b);
      mc_st(mc + i + 2 * stride, c);
      mc_st(mc + i + 3 * stride, d);
    }
    for (; i < s1; i += stride) mc_st(mc + i, mc_ld_reduce(mc + i));
  } else {
    for (size_t i = s0 + (size_t)blockIdx.x * kThreads + threadIdx.x; i < s1; i += stride) {
      float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      float4 v[kMaxWorld];
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < P.world) v[p] = ld_peer(reinterpret_cast<const float4*>(P.bufs[p]) + i);
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)                      // fixed order: the same sum on every rank
        if (p < P.world) { acc.x += v[p].x; acc.y += v[p].y; acc.z += v[p].z; acc.w += v[p].w; }
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < P.world) reinterpret_cast<float4*>(P.bufs[p])[i] = acc;
    }
  }
  peer_barrier(P, 1, epoch);
  if (threadIdx.x == 0) {
    if (atomicAdd(P.epoch + 1, 1u) == gridDim.x - 1) {
      P.epoch[1] = 0u;
      __threadfence();
      P.epoch[0] = epoch;
    }
  }
}

}  // namespace ar
}  // namespace tvm

using namespace tvm;

extern "C" int tvm_allreduce_signal_words(int world, size_t* out_words) {
  TVM_REQUIRE(out_words && world >= 1 && world <= ar::kMaxWorld, "bad arguments");
  *out_words = (size_t)2 * TVM_AR_MAX_CTAS * world;
  return 0;
}

extern "C" int tvm_allreduce_sum(const TvmPeerComm* c, size_t offset_floats, size_t n_floats, int n_ctas, void* stream_) {
  TVM_REQUIRE(c && c->world >= 1 && c->world <= ar::kMaxWorld && c->rank >= 0 && c->rank < c->world, "bad communicator");
  TVM_REQUIRE(c->epoch_dev != nullptr, "null epoch counter");
  TVM_REQUIRE((offset_floats & 3) == 0 && (n_floats & 3) == 0, "offset and length must be multiples of 4 floats");
  TVM_REQUIRE(n_ctas >= 1 && n_ctas <= TVM_AR_MAX_CTAS, "n_ctas must be in 1..%d", TVM_AR_MAX_CTAS);
  if (c->world == 1 || n_floats == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  ar::Params P;
  for (int p = 0; p < c->world; ++p) {
    TVM_REQUIRE(c->bufs[p] && c->signals[p], "null peer pointer");
    TVM_REQUIRE(((uintptr_t)c->bufs[p] & 15) == 0, "peer buffers must be 16-byte aligned");
    P.bufs[p] = (float*)c->bufs[p];
    P.signals[p] = c->signals[p];
  }
  P.multicast = (float*)c->multicast;
  P.epoch = c->epoch_dev;
  P.rank = c->rank;
  P.world = c->world;
  P.offset4 = offset_floats / 4;
  P.n4 = n_floats / 4;
  if (P.multicast) ar::k_allreduce<true><<<n_ctas, ar::kThreads, 0, stream>>>(P);
  else ar::k_allreduce<false><<<n_ctas, ar::kThreads, 0, stream>>>(P);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
