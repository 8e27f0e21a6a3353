// What slows tcgen05.mma down inside the conv kernel?  (conv_tc.cu measures ~140 clk per M128 x N72 x K8 SS-mode
// MMA against 46 clk for the same instruction issued back to back in isolation.)
//   ./umma_bench2 <mode bits> [N]
// One CTA of 544 threads.  Warp 8 / lane 0 issues 64 rounds of 27 accumulating MMAs over 8 distinct A tiles
// and 3 B slabs (like one y step of the 8->8 3D conv) and reports cycles per MMA, while, depending on the mode
// bits, the other warps generate the traffic the real kernel has:
//   1: warps 0-7 read the OTHER accumulator buffer from TMEM with tcgen05.ld 32x32b.x4 (3 loads + wait, like the epilogue)
//   2: same but 32x32b.x16 loads (same bytes, 4x fewer instructions)
//   4: warps 9-16 stream 16-byte shared-memory stores into an operand-ring sized region (like the producers)
//   8: warps 9-16 stream global loads (L2 hits)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}

__global__ void __launch_bounds__(544, 1) bench(int mode, int N, const float* gsrc, float* sink, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // smem: [0, 64 KB) A tiles (8 x 8 KB), [64 KB, 64 KB + 3 * N * 64) B slabs, [96 KB, 200 KB) store target
  for (int i = tid; i < (96 << 10) / 4; i += 544) reinterpret_cast<float*>(smem)[i] = (float)(i & 1023) * 1e-3f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    stop = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  float acc = 0.f;
  if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t a0 = make_desc(smem_u32(smem), 2048, 128), b0 = make_desc(smem_u32(smem + (64 << 10)), N * 16, 128);
      const long long t0 = clock64();
      const int rounds = 64;
      for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int zo = 0; zo < 3; ++zo)
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const uint64_t a = a0 + (uint64_t)((((r + zo + kd) & 7) * 8192) >> 4);
            const uint64_t b = b0 + (uint64_t)((kd * N * 64) >> 4);
            const uint32_t d = tmem + zo * N;
            mma_tf32(d, a, b, idesc, kd > 0);
            mma_tf32(d, a + (4096 >> 4), b, idesc, 1);
            mma_tf32(d, a, b + ((N * 32) >> 4), idesc, 1);
          }
      }
      commit(&bar);
      wait_bar(&bar, 0);
      cycles[0] = (clock64() - t0) / (rounds * 27);
      stop = 1;
    }
  } else if (warp < 8 && (mode & 3)) {
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 4;
    while (!stop) {
      if (mode & 1) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          uint32_t u[12];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%12];\n\t"
              "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4,%5,%6,%7}, [%13];\n\t"
              "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8,%9,%10,%11}, [%14];\n\t"
              "tcgen05.wait::ld.sync.aligned;"
              : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                "=r"(u[10]), "=r"(u[11])
              : "r"(tb + i * 24), "r"(tb + i * 24 + 8), "r"(tb + i * 24 + 16)
              : "memory");
#pragma unroll
          for (int j = 0; j < 12; ++j) acc += __uint_as_float(u[j]);
        }
      } else {
        uint32_t u[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
            "tcgen05.wait::ld.sync.aligned;"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
              "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
            : "r"(tb)
            : "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += __uint_as_float(u[j]);
      }
      // ~ the arithmetic the epilogue does between loads
#pragma unroll 8
      for (int j = 0; j < 64; ++j) acc = fmaf(acc, 1.0001f, 0.5f);
    }
  } else if (warp >= 9 && (mode & 12)) {
    const int m = (warp - 9) * 32 + lane;
    int it = 0;
    while (!stop) {
      if (mode & 8) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 20; ++j) v += __ldg(gsrc + ((it * 20 + j) * 4096 + m * 4) % (1 << 22));
        acc += v;
      }
      if (mode & 4) {
        uint8_t* sb = smem + (96 << 10) + (it & 1) * (40 << 10) + (m & 127) * 16 + (m >> 7) * 2048;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          *reinterpret_cast<float4*>(sb + j * 8192) = make_float4(acc, acc, acc, acc);
          *reinterpret_cast<float4*>(sb + j * 8192 + 4096) = make_float4(acc, acc, acc, acc);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
#pragma unroll 8
      for (int j = 0; j < 40; ++j) acc = fmaf(acc, 1.0001f, 0.5f);
      ++it;
    }
  }
  if (acc == 12345.678f) sink[tid] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main(int argc, char** argv) {
  const int N = argc > 2 ? atoi(argv[2]) : 72;
  float *g, *sink;
  long long* c;
  cudaMalloc(&g, (size_t)(1 << 22) * 4 + 4096);
  cudaMemset(g, 0, (size_t)(1 << 22) * 4 + 4096);
  cudaMalloc(&sink, 4096);
  cudaMalloc(&c, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
  const char* names[16] = {"MMA alone", "+ tmem.ld x4", "+ tmem.ld x16", "", "+ smem stores", "+ ld x4 + sts", "+ ld x16 + sts", "",
                           "+ global loads", "+ ld x4 + ldg", "+ ld x16 + ldg", "", "+ sts + ldg", "+ ld x4 + sts + ldg", "+ ld x16 + sts + ldg", ""};
  const int modes[] = {0, 1, 2, 4, 8, 5, 13, 14};
  for (int mi = 0; mi < 8; ++mi) {
    const int mode = argc > 1 && atoi(argv[1]) >= 0 ? atoi(argv[1]) : modes[mi];
    bench<<<1, 544, 200 << 10>>>(mode, N, g, sink, c);
    cudaError_t e = cudaDeviceSynchronize();
    long long cy = 0;
    cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    printf("N=%d mode %2d (%-22s): %lld cycles per MMA  [%s]\n", N, mode, names[mode], cy, cudaGetErrorString(e));
    if (argc > 1 && atoi(argv[1]) >= 0) break;
  }
  return 0;
}
