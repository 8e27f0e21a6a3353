// Shared device/host helpers for libesm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/esm_b200.h"

namespace esm {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define ESM_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      esm::set_error(__VA_ARGS__);        \
      return ESM_ERR_ARG;                 \
    }                                     \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// Packed fp32x2 FMA (Blackwell FFMA2): d = a * b + d on both lanes.  ptxas folds a {x,x} pair into
// the scalar-broadcast operand form, so the broadcast costs no extra instruction.  Measured on B200:
// same FMA rate as scalar FFMA at half the issue slots (scratch/fma_bench.cu), which is what lets the
// shared-memory loads of the direct convolution hide under the math.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ESM_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
    case ESM_ACT_RELU: return fmaxf(x, 0.0f);
    case ESM_ACT_SILU: return x / (1.0f + expf(-x));
    case ESM_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case ESM_ACT_2SIGMOID: return 2.0f * (1.0f / (1.0f + expf(-x)));
    case ESM_ACT_RELU6: return fminf(fmaxf(x, 0.0f), 6.0f);
    default: return x;
  }
}

// Activation of 4 values at once, kept out of line on purpose: one copy of the erff / expf code per
// kernel instead of one per call site, with 4 independent evaluations in flight per call.
static __device__ __noinline__ float4 apply_act4(float4 v, int act) {
  switch (act) {
#define ESM_ACT4_CASE(CODE)                                                                                   \
  case CODE:                                                                                                  \
    return make_float4(apply_act(v.x, CODE), apply_act(v.y, CODE), apply_act(v.z, CODE), apply_act(v.w, CODE));
    ESM_ACT4_CASE(ESM_ACT_GELU)
    ESM_ACT4_CASE(ESM_ACT_RELU)
    ESM_ACT4_CASE(ESM_ACT_SILU)
    ESM_ACT4_CASE(ESM_ACT_SIGMOID)
    ESM_ACT4_CASE(ESM_ACT_2SIGMOID)
    ESM_ACT4_CASE(ESM_ACT_RELU6)
#undef ESM_ACT4_CASE
    default:
      return v;
  }
}

__device__ __forceinline__ int ceil_div_dev(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

}  // namespace esm
