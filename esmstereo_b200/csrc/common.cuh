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

// Programmatic dependent launch (PDL).  A kernel launched through launch_k(pdl = true, ...) may start while the previous
// kernel of the stream is still draining: its CTAs are placed as soon as every CTA of that kernel has executed
// pdl_launch_dependents() (first instruction of every kernel here) and resources are free.  EVERY kernel launched that
// way executes pdl_wait() in every thread before its first global-memory access (weights included), which blocks until
// the previous grid has completed and flushed; what overlaps is launch latency, CTA placement and the on-chip prologue
// (mbarrier init, TMEM allocation).  ESM_PDL is a bit mask of the kernel families that use it: 1 FP32-pipe conv, 16
// pointwise, 32 everything else on the forward path.  The tcgen05 kernels (bits 2 / 4 / 8 in the round-2 experiment) do
// NOT carry the instructions any more: a `griddepcontrol.wait` in their prologue makes ptxas give up the uniform datapath
// for the values computed around it (R2UR 20 -> 380 in the gwc stem kernel, +14 us; the MMA issue loop pays 11 R2URs per
// stage), which cost more than the launch overlap ever returned.
bool pdl_enabled(int family_bit);
#ifdef ESM_NO_PDL
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

template <typename... Params, typename... Args>
static inline cudaError_t launch_k(bool pdl, void (*fn)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, fn, static_cast<Params>(args)...);
}

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

// torch upsample_bilinear2d, align_corners=False, scale 1 / rscale, of one [h, w] plane at output pixel (Y, X)
// (src = (dst + 0.5) * rscale - 0.5, clamped at 0).  Explicit roundings: esm_bilinear_add_f32 and the PixelShuffle
// epilogue of the conv engine (the final assembly fused into conv1_up's sub-pixel form) must agree bit for bit, whatever
// the compiler would contract in either context.
__device__ __forceinline__ float bilinear_up(const float* __restrict__ pb, int h, int w, int Y, int X, float rscale) {
  float sy = __fmaf_rn((float)Y + 0.5f, rscale, -0.5f);
  float sx = __fmaf_rn((float)X + 0.5f, rscale, -0.5f);
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + ((y0 < h - 1) ? 1 : 0), x1 = x0 + ((x0 < w - 1) ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float v00 = __ldg(pb + (long long)y0 * w + x0), v01 = __ldg(pb + (long long)y0 * w + x1);
  const float v10 = __ldg(pb + (long long)y1 * w + x0), v11 = __ldg(pb + (long long)y1 * w + x1);
  const float top = __fmaf_rn(lx, v01, __fmul_rn(hx, v00));
  const float bot = __fmaf_rn(lx, v11, __fmul_rn(hx, v10));
  return __fmaf_rn(ly, bot, __fmul_rn(hy, top));
}

__device__ __forceinline__ int ceil_div_dev(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

}  // namespace esm
