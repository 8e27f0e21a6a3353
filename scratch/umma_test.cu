// Standalone probe of tcgen05.mma kind::tf32 (SS mode, K-major, no swizzle) on sm_100a:
//   ./umma_test <N> [reps]
// checks D = A(128x8) * B(Nx8)^T against the CPU for (a) tf32-exact inputs, (b) full fp32 inputs under
// the truncation and the round-to-nearest hypothesis, (c) the 3-pass hi/lo split; then times `reps`
// back-to-back MMAs (cycles per dispatch, shared-memory operand read included).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout type 0 = no swizzle
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
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

// A: [2][128][4] floats (k-half, row, 4 k), B: [2][N][4]
__global__ void __launch_bounds__(128, 1) umma_kernel(const float* A, const float* B, float* D, int N, int npass, int reps, long long* cycles) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float* sA = smem;                       // npass x 4 KB
  float* sB = smem + npass * 1024;        // npass x N*8 floats
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < npass * 1024; i += 128) sA[i] = A[i];
  for (int i = tid; i < npass * N * 8; i += 128) sB[i] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D=f32, A=B=tf32, K-major both, N>>3 at bit 17, M>>4 at bit 24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    // pass p: A_p x B_p, all accumulated (caller arranges hi/lo operands per pass)
    for (int p = 0; p < npass; ++p) {
      const uint64_t da = make_desc(smem_u32(sA + p * 1024), 128 * 16, 128);
      const uint64_t db = make_desc(smem_u32(sB + p * N * 8), N * 16, 128);
      mma_tf32(tmem, da, db, idesc, p > 0);
    }
    commit(&bar);
  }
  wait_bar(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read back: warp w owns lanes 32w..32w+31; 8 columns at a time
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
  // timing: reps back-to-back MMAs on the same operands
  if (reps > 0) {
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
      const uint64_t da = make_desc(smem_u32(sA), 128 * 16, 128);
      const uint64_t db = make_desc(smem_u32(sB), N * 16, 128);
      t0 = clock64();
      for (int i = 0; i < reps; ++i) mma_tf32(tmem + 256, da, db, idesc, i > 0);
      commit(&bar);
    }
    wait_bar(&bar, 1);
    if (tid == 0) {
      t1 = clock64();
      cycles[0] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static float trunc_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
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
  const int N = argc > 1 ? atoi(argv[1]) : 32;
  const int reps = argc > 2 ? atoi(argv[2]) : 512;
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  std::vector<float> a(128 * 8), b(N * 8);
  for (auto& v : a) v = rnd();
  for (auto& v : b) v = rnd();
  a[0] = 1.f + ldexpf(1.f, -11) + ldexpf(1.f, -12);  // trunc -> 1, RN -> 1 + 2^-10
  for (int k = 1; k < 8; ++k) a[k] = 0.f;
  b[0] = 1.f;
  auto pack = [](const std::vector<float>& m, int rows, std::vector<float>* out, size_t off) {
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < 8; ++k) (*out)[off + (size_t)(k / 4) * rows * 4 + r * 4 + (k % 4)] = m[r * 8 + k];
  };
  float *dA, *dB, *dD;
  long long* dC;
  cudaMalloc(&dA, 3 * 1024 * 4);
  cudaMalloc(&dB, 3 * N * 8 * 4);
  cudaMalloc(&dD, 128 * N * 4);
  cudaMalloc(&dC, 8);
  std::vector<float> d(128 * N);
  auto run = [&](int npass, const std::vector<float>& pa, const std::vector<float>& pb, int r) {
    cudaMemcpy(dA, pa.data(), pa.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, pb.data(), pb.size() * 4, cudaMemcpyHostToDevice);
    const size_t sm = (size_t)npass * (4096 + N * 32) + 1024;
    cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    umma_kernel<<<1, 128, sm>>>(dA, dB, dD, N, npass, r, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("N=%d: launch failed: %s\n", N, cudaGetErrorString(e));
      exit(1);
    }
    cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost);
  };
  // (1) single pass on raw fp32 operands
  std::vector<float> pa(1024), pb(N * 8);
  pack(a, 128, &pa, 0);
  pack(b, N, &pb, 0);
  run(1, pa, pb, reps);
  long long cyc = 0;
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  double et = 0, er = 0, ef = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double st = 0, sr = 0, sf = 0;
      for (int k = 0; k < 8; ++k) {
        st += (double)trunc_tf32(a[m * 8 + k]) * trunc_tf32(b[n * 8 + k]);
        sr += (double)rn_tf32(a[m * 8 + k]) * rn_tf32(b[n * 8 + k]);
        sf += (double)a[m * 8 + k] * b[n * 8 + k];
      }
      et = fmax(et, fabs(d[m * N + n] - st));
      er = fmax(er, fabs(d[m * N + n] - sr));
      ef = fmax(ef, fabs(d[m * N + n] - sf));
    }
  printf("N=%d 1-pass: max|d-trunc|=%.3e max|d-rn|=%.3e max|d-fp32|=%.3e  D[0][0]=%.10f (trunc->1.0, rn->1.0009765625)\n", N, et, er, ef, d[0]);
  printf("N=%d timing: %d MMAs in %lld cycles = %.1f cyc/MMA (floor N/2 = %d)\n", N, reps, cyc, (double)cyc / reps, N / 2);
  // (2) 3-pass split: hi = trunc(x), lo = x - hi;  hi*hi + lo*hi + hi*lo
  std::vector<float> ah(a), al(a), bh(b), bl(b);
  for (size_t i = 0; i < a.size(); ++i) { ah[i] = trunc_tf32(a[i]); al[i] = a[i] - ah[i]; }
  for (size_t i = 0; i < b.size(); ++i) { bh[i] = trunc_tf32(b[i]); bl[i] = b[i] - bh[i]; }
  std::vector<float> pa3(3 * 1024), pb3(3 * N * 8);
  pack(a, 128, &pa3, 0);      // raw fp32: the hardware is expected to see hi
  pack(bh, N, &pb3, 0);
  pack(al, 128, &pa3, 1024);
  pack(bh, N, &pb3, (size_t)N * 8);
  pack(a, 128, &pa3, 2048);
  pack(bl, N, &pb3, (size_t)2 * N * 8);
  run(3, pa3, pb3, 0);
  double e3 = 0, ref_max = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double sf = 0;
      for (int k = 0; k < 8; ++k) sf += (double)a[m * 8 + k] * b[n * 8 + k];
      e3 = fmax(e3, fabs(d[m * N + n] - sf));
      ref_max = fmax(ref_max, fabs(sf));
    }
  printf("N=%d 3-pass split: max|d-fp64|=%.3e (|ref| up to %.2f; fp32 eps*|ref| = %.1e)\n", N, e3, ref_max, ref_max * 6e-8);
  return 0;
}
