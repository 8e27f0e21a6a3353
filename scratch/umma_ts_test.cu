// Probe of tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (TS form): A[m][k] at TMEM lane m, column a0 + k
// (written with tcgen05.st.32x32b.x8 by the thread that owns lane m), B in shared memory (K-major, no swizzle).
//   ./umma_ts_test <N> [reps]
// checks D = A(128x8) * B(Nx8)^T against the CPU on tf32-exact inputs, then times `reps` back-to-back MMAs in TS and SS form.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

// A: [128][8] row-major floats; B: [2][N][4] (k-half, row, 4 k)
__global__ void __launch_bounds__(128, 1) ts_kernel(const float* A, const float* B, float* D, int N, int reps, long long* cycles) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float* sB = smem;          // N*8 floats
  float* sA = smem + N * 8;  // SS-form copy of A for the timing comparison: [2][128][4]
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < N * 8; i += 128) sB[i] = B[i];
  for (int i = tid; i < 1024; i += 128) {
    const int kh = i / 512, m = (i / 4) % 128, k = i % 4;
    sA[i] = A[m * 8 + kh * 4 + k];
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_col = 384;  // A tile: columns 384..391
  // every thread stores its row of A: lane = 32 * warp + laneid, 8 consecutive columns
  {
    uint32_t r[8];
    for (int k = 0; k < 8; ++k) r[k] = __float_as_uint(A[tid * 8 + k]);
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + a_col;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
                 "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t db = make_desc(smem_u32(sB), N * 16, 128);
  if (tid == 0) {
    mma_ts(tmem, tmem + a_col, db, idesc, 0);
    commit(&bar);
  }
  wait_bar(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  uint32_t par = 1;
  for (int form = 0; form < 2; ++form) {  // 0: TS, 1: SS
    long long t0 = 0;
    if (tid == 0) {
      const uint64_t da = make_desc(smem_u32(sA), 128 * 16, 128);
      t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        if (form == 0)
          mma_ts(tmem + 128, tmem + a_col, db, idesc, i > 0);
        else
          mma_ss(tmem + 128, da, db, idesc, i > 0);
      }
      commit(&bar);
    }
    wait_bar(&bar, par);
    par ^= 1;
    if (tid == 0) cycles[form] = clock64() - t0;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static float rn_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 40;
  const int reps = argc > 2 ? atoi(argv[2]) : 512;
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  std::vector<float> a(128 * 8), b(N * 8), bp(N * 8);
  for (auto& v : a) v = rn_tf32(rnd());
  for (auto& v : b) v = rn_tf32(rnd());  // b[n][k]
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 8; ++k) bp[(k / 4) * N * 4 + n * 4 + (k % 4)] = b[n * 8 + k];
  float *dA, *dB, *dD;
  long long* dC;
  cudaMalloc(&dA, a.size() * 4);
  cudaMalloc(&dB, bp.size() * 4);
  cudaMalloc(&dD, 128 * N * 4);
  cudaMalloc(&dC, 16);
  cudaMemcpy(dA, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (N * 8 + 1024) * 4 + 1024;
  cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ts_kernel<<<1, 128, smem>>>(dA, dB, dD, N, reps, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> d(128 * N);
  long long cyc[2];
  cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 8; ++k) s += (double)a[m * 8 + k] * b[n * 8 + k];
      maxerr = fmax(maxerr, fabs(s - d[m * N + n]));
    }
  printf("N=%d TS-form max |err| vs fp64 on tf32-exact inputs: %.3g  (D[0][0]=%g D[127][N-1]=%g)\n", N, maxerr, d[0], d[127 * N + N - 1]);
  printf("N=%d  %d back-to-back MMAs: TS %.1f clk/MMA, SS %.1f clk/MMA\n", N, reps, (double)cyc[0] / reps, (double)cyc[1] / reps);
  return 0;
}
