// Does a UMMA A operand whose start address is moved by whole rows (the kw / kh taps of the flat conv engine) run at
// full rate, and is it read correctly?  Layouts: K-major no-swizzle (rows 16 B apart), SWIZZLE_32B / 64B / 128B.
//   ./umma_shift
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__host__ __device__ inline float Aval(int r, int k) { return (float)((r * 3 + k * 5) % 17 - 8); }
__host__ __device__ inline float Bval(int n, int k) { return (float)((n * 7 + k) % 13 - 6); }

// mode: 0 none, 1 SW32, 2 SW64, 3 SW128.  Row pitch 16 / 32 / 64 / 128 bytes; KTOT = pitch / 4 channels per row (mode 0: 8).
__global__ void __launch_bounds__(160, 1) bench(int mode, int shift, int use_bo, int N, int kofs, float* dout, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ROWS = 288;
  uint8_t* sA = smem;              // 1024-aligned
  uint8_t* sB = smem + (48 << 10);
  const int pitch = mode == 0 ? 16 : (16 << mode);
  const int ktot = mode == 0 ? 8 : pitch / 4;
  for (int i = tid; i < ROWS * ktot; i += 160) {
    const int r = i / ktot, k = i % ktot;
    uint32_t off;
    if (mode == 0) {
      off = (k / 4) * (ROWS * 16) + r * 16 + (k % 4) * 4;
    } else {
      const uint32_t lin = r * pitch + k * 4;
      const uint32_t bits = mode == 1 ? 1 : mode == 2 ? 3 : 7;
      off = lin ^ (((lin >> 7) & bits) << 4);
    }
    *reinterpret_cast<float*>(sA + off) = Aval(r, k);
  }
  for (int i = tid; i < N * 8; i += 160) {
    const int n = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sB + (k / 4) * (N * 16) + n * 16 + (k % 4) * 4) = Bval(n, k);
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_addr = smem_u32(sA) + shift * pitch + (mode == 0 ? 0 : kofs * 4);
    uint64_t da;
    if (mode == 0) {
      da = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)(((ROWS * 16) >> 4) & 0x3FFF) << 16) | ((uint64_t)((128 >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
    } else {
      const uint64_t lt = mode == 1 ? 6 : mode == 2 ? 4 : 2;
      const uint32_t sbo = 8 * pitch;
      const uint64_t bo = use_bo ? ((a_addr >> 7) & 7) : 0;
      da = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | (bo << 49) | (lt << 61);
    }
    const uint64_t db = (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF) | ((uint64_t)(((N * 16) >> 4) & 0x3FFF) << 16) | ((uint64_t)((128 >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
    // correctness: one non-accumulating MMA into columns [0, N)
    mma_tf32(tmem, da, db, idesc, 0);
    commit(&bar);
    wait_bar(&bar, 0);
    // timing: 512 accumulating MMAs into columns [256, 256 + N)
    const long long t0 = clock64();
    for (int i = 0; i < 512; ++i) mma_tf32(tmem + 256, da, db, idesc, i > 0);
    commit(&bar);
    wait_bar(&bar, 1);
    cycles[0] = (clock64() - t0) / 512;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    for (int c = 0; c < N; c += 8) {
      uint32_t u[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                   : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c)
                   : "memory");
      for (int j = 0; j < 8; ++j) dout[(warp * 32 + lane) * N + c + j] = __uint_as_float(u[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  const int N = 64;
  float* d;
  long long* c;
  cudaMalloc(&d, 128 * N * 4);
  cudaMalloc(&c, 8);
  float* h = (float*)malloc(128 * N * 4);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10);
  const char* names[4] = {"no-swizzle", "SW32", "SW64", "SW128"};
  for (int mode = 0; mode < 4; ++mode)
    for (int bo = 0; bo < (mode ? 2 : 1); ++bo)
      for (int shift : {0, 1, 2, 3, 4, 8, 9}) {
        const int kofs = mode >= 2 ? 8 : 0;  // swizzled wide rows: use the second K-slice to exercise the K advance too
        bench<<<1, 160, 100 << 10>>>(mode, shift, bo, N, kofs, d, c);
        cudaError_t e = cudaDeviceSynchronize();
        long long cy = 0;
        cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(h, d, 128 * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            float want = 0.f;
            for (int k = 0; k < 8; ++k) want += Aval(m + shift, kofs + k) * Bval(n, k);
            if (h[m * N + n] != want) ++bad;
          }
        printf("%-10s base_offset %d shift %d rows: %4lld clk/MMA (M128 N%d K8)  %s  [%s]\n", names[mode], bo, shift, cy, N,
               bad ? "WRONG" : "exact", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
