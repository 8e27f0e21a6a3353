// Launch-overhead microbench (diagnostic): a chain of N dependent persistent-style kernels (148 CTAs x 608 threads,
// ~200 KB of dynamic shared memory, 512 TMEM columns -- the footprint of the tcgen05 conv engines) replayed from a
// CUDA graph, with and without programmatic dependent launch.  Tells how much of a ~15 us small layer is
// launch gap + CTA ramp + prologue, and how much of it PDL can hide.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/pdl_bench scratch/pdl_bench.cu && scratch/pdl_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode bit 0: TMEM alloc / dealloc; bit 1: PDL instructions; spin_ns: busy time per CTA (+ imbalance by blockIdx)
__global__ void __launch_bounds__(608, 1) chain_kernel(int mode, int spin_ns, int skew_ns, const float* in, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem);
  if (mode & 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (mode & 1) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // a prologue that does not depend on the previous kernel (weight staging): ~1 us of L2 reads
  float acc = 0.f;
  if (mode & 4) {
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) acc += __ldg(in + i);
    reinterpret_cast<float*>(smem)[64 + threadIdx.x] = acc;
    __syncthreads();
  }
  if (mode & 2) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (spin_ns > 0) {
    const long long t0 = clock64();
    const long long clk = (long long)(spin_ns + (blockIdx.x % 8) * skew_ns) * 19 / 10;  // ~1.9 GHz
    while (clock64() - t0 < clk) {}
  }
  if (threadIdx.x == 0) out[blockIdx.x] = in[blockIdx.x] + acc;
  if (mode & 1) {
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*slot), "r"(512) : "memory");
  }
}

static float run(int n, int grid, int threads, int smem, int mode, int spin, int skew, float* a, float* b) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaFuncSetAttribute((const void*)chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (mode & 2) ? 1 : 0;
    const float* in = (i & 1) ? b : a;
    float* out = (i & 1) ? a : b;
    cudaLaunchKernelEx(&cfg, chain_kernel, mode, spin, skew, in, out);
  }
  cudaStreamEndCapture(st, &g);
  if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) {
    printf("instantiate failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return -1.f;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("error: %s\n", cudaGetErrorString(err));
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaStreamDestroy(st);
  return ms * 1e3f / (10.f * n);
}

int main() {
  float *a, *b;
  cudaMalloc(&a, 1 << 20);
  cudaMalloc(&b, 1 << 20);
  cudaMemset(a, 0, 1 << 20);
  cudaMemset(b, 0, 1 << 20);
  const int n = 100;
  struct Cfg { const char* name; int grid, threads, smem; } cfgs[] = {
      {"148 x 608 thr, 200 KB smem", 148, 608, 200 << 10},
      {"148 x 608 thr, 100 KB smem", 148, 608, 100 << 10},
      {"296 x 256 thr,  64 KB smem", 296, 256, 64 << 10},
      {"936 x 128 thr,   0 KB smem", 936, 128, 0},
  };
  for (const Cfg& c : cfgs) {
    printf("== %s ==\n", c.name);
    for (int spin : {0, 5000, 15000}) {
      for (int skew : {0, 300}) {
        if (spin == 0 && skew) continue;
        for (int tm = 0; tm < 2; ++tm) {
          if (tm && c.threads != 608) continue;
          const int base = (tm ? 1 : 0) | 4;
          const float t0 = run(n, c.grid, c.threads, c.smem, base, spin, skew, a, b);
          const float t1 = run(n, c.grid, c.threads, c.smem, base | 2, spin, skew, a, b);
          printf("spin %5d ns skew %3d tmem %d : plain %6.2f us/kernel   pdl %6.2f us/kernel\n", spin, skew, tm, t0, t1);
        }
      }
    }
  }
  return 0;
}
