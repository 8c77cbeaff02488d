// Shared declarations of the backward kernels (tvm_backward.cu: fp32; tvm_bwd_tc.cu: tcgen05).
#pragma once
#include "tvm_common.cuh"

namespace tvm {

struct BwdParams {
  FwdParams f;
  const float* d_rgb_map;
  const float* d_penalty;     // TVM_VARIANT_REF: device scalar dL/d penalty, or NULL
  TvmGrads g;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace tvm
