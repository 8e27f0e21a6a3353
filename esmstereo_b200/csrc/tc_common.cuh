// PTX wrappers shared by the tcgen05 convolution kernels (conv_tc.cu: taps in N, resident weights;
// conv_tcg.cu: taps in K, streamed weights): mbarriers, UMMA descriptors, tcgen05.mma / commit / ld,
// TF32 rounding and the epilogue GELU.
#pragma once
#include "common.cuh"

namespace esm {

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// Wait on an mbarrier phase.  A watchdog turns a pipeline deadlock (a bug) into a trapped launch ("unspecified launch
// failure") instead of a hung GPU: ~2 s of failed polls, far beyond any legitimate wait here.
#ifndef TC_DEADLOCK_PRINTF
// The watchdog traps without a message: a printf here is an ABI call (vprintf) inside every wait loop, and a kernel that
// contains one is register-allocated around it -- A/B of two builds: 1.952 -> 1.916 ms per KITTI pair without the call
// (and ptxas held the regions of a setmaxnreg experiment to the smallest budget of the kernel).  -DTC_DEADLOCK_PRINTF
// (ESM_AB_DEFS, esmstereo_b200/build.py) brings the message back for debugging a pipeline.
static __device__ __forceinline__ void tc_deadlock(int, uint32_t) { __trap(); }
#else
static __device__ __noinline__ void tc_deadlock(int tag, uint32_t parity) {
  if ((threadIdx.x & 31) == 0)
    printf("esm tc_conv: deadlock in block %d warp %d waiting on barrier %d parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), tag, parity);
  __trap();
}
#endif
#ifndef TC_WAIT_HINT_NS
#define TC_WAIT_HINT_NS 20000  // mbarrier.try_wait suspend-time hint in ns (0: the system default, measured ~100 clk per poll; A/B: 1.9723 -> 1.9633 ms per KITTI pair)
#endif
#ifdef TC_PROFILE
// role profiler (build with ESM_TC_PROFILE=1): cycles block 0's warps spend in each class of mbarrier wait
static __device__ unsigned long long tc_prof_wait[32][8];
#endif
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
#ifdef TC_PROFILE
  const long long t_begin = clock64();
#endif
  const uint32_t addr = tc_smem_u32(bar);
  uint32_t done, polls = 0;
  do {
#if TC_WAIT_HINT_NS > 0
    // suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires) instead of
    // re-polling every ~100 clk -- the polls of the idle roles were half of the instructions these kernels issued
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(addr), "r"(parity), "r"((uint32_t)TC_WAIT_HINT_NS)
                 : "memory");
    if (!done && ++polls > (2000000000u / (uint32_t)TC_WAIT_HINT_NS)) tc_deadlock(tag, parity);
#else
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
    if (!done && ++polls > (1u << 23)) tc_deadlock(tag, parity);
#endif
  } while (!done);
#ifdef TC_PROFILE
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) tc_prof_wait[threadIdx.x >> 5][tag / 100] += (unsigned long long)(clock64() - t_begin);
#endif
}
// K-major, no-swizzle UMMA shared-memory descriptor: 8-row x 16-byte core matrices, rows 16 bytes
// apart; LBO = byte distance between the two K halves, SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// TS form: the A operand (128 rows x 8 tf32) comes from tensor memory -- row m at TMEM lane m, k at column a + k
// (scratch/umma_ts_test.cu) -- so it costs no shared-memory bandwidth; B stays a shared-memory descriptor.
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// registers -> TMEM: this warp's lane quadrant, 4 consecutive columns (asynchronous: tc_st_wait() before signalling)
__device__ __forceinline__ void tc_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
               "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
               : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// TMEM -> registers: this warp's lane quadrant, 16 / 4 consecutive columns.  Asynchronous: tc_ld_wait()
// before the first use (all three are volatile with a memory clobber, so they keep their order).
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* r) {
  uint32_t u[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                 "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, float* r) {
  uint32_t u[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float* r) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
// mbarrier arrive that also announces `bytes` of asynchronous (TMA) traffic to come
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
// 1D bulk copy global -> shared (TMA, no tensor map): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(tc_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t tc_elect() {
  uint32_t leader;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(leader));
  return leader;
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Split helpers.  hi = x rounded to the nearest TF32 (ties away, low 13 mantissa bits zero) with two full-rate integer
// instructions: `cvt.rna.tf32.f32` is emulated with four on sm_100a, and the native `cvt.rn.tf32.f32` (F2FP.TF32) runs on
// a narrow conversion pipe -- ncu showed math-pipe throttling in the producers, which convert 72 values per stage.
// lo = x - hi needs no rounding of its own: the tensor core truncates its operands to TF32 (scratch/umma_test.cu), and
// truncating lo costs 2^-21 |x| at most, unbiased (lo has either sign).
__device__ __forceinline__ float tc_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float tc_lo(float x, float hi) { return x - hi; }

// GELU for the epilogue warps, which bound most of these kernels: erff() costs ~35 instructions per value on a
// divergent warp (two branches), this one 16, branch-free.  erfc(t) = 2^p(t) with p a degree-8 fit of
// log2(erfcx(t)) - t^2 log2(e) on [0, 4] weighted by the GELU's sensitivity, so that
//   gelu(x) = x - h (x >= 0),  h (x < 0),   h = 0.5 x erfc(|x| / sqrt 2).
// Absolute error <= 6e-8 (an ulp of an O(1) activation), relative error <= 3e-7 for x >= 0; the fit and its
// error table are in tests/test_host_cpu.py::test_tc_gelu_polynomial.
// Mean relative truncation error of one tcgen05 accumulate (the tensor core rounds its fp32 accumulator toward zero),
// measured with scripts/engine_accuracy.py on coherent sums; the epilogues scale an accumulator that chained n
// accumulates by 1 + n * TC_TRUNC_BIAS.
constexpr float TC_TRUNC_BIAS = 1.25e-8f;

// SiLU for the same epilogues: x / (1 + 2^(-x log2 e)) with ex2.approx and rcp.approx (relative error ~3e-7); expf + an
// IEEE division cost the FP32-pipe epilogue 21 us on the 16 -> 64 UpShuffle layer.
__device__ __forceinline__ float tc_silu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
__device__ __forceinline__ float tc_gelu(float x) {
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float q = -2.906944503e-05f;
  q = fmaf(q, t, 3.042682386e-04f);
  q = fmaf(q, t, -1.000199492e-03f);
  q = fmaf(q, t, -1.645459926e-03f);
  q = fmaf(q, t, 2.910655108e-02f);
  q = fmaf(q, t, -1.489377188e-01f);
  q = fmaf(q, t, -9.182927772e-01f);
  q = fmaf(q, t, -1.627922676e+00f);
  q = fmaf(q, t, 3.958853834e-07f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  const float h = (0.5f * x) * e;
  return x >= 0.f ? x - h : h;
}

}  // namespace esm
