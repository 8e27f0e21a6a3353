// Measured dense TF32 tensor-core peak of this device: every SM issues back-to-back M128 x N256 x K8 tcgen05.mma
// (kind::tf32, SS form, accumulators in TMEM) from one elected thread.  This is the denominator bench.py holds the
// conv engines' tensor rooflines against (MEASURED_PEAKS.json only has the bf16 figure).
#include "tc_common.cuh"

namespace esm {

__global__ void __launch_bounds__(160, 1) umma_tf32_peak_kernel(int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int N = 256;
  // A: [2 K-halves][128 rows][16 B] at 0 (LBO 2048); B: [2][256 rows][16 B] at 8 KB (LBO 4096)
  for (int i = tid; i < (16 << 10) / 4; i += 160) reinterpret_cast<float*>(smem)[i] = 0.5f + (float)(i & 255) * (1.0f / 512.0f);
  if (tid == 0) tc_mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (warp == 4 && tc_elect()) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = tc_desc(tc_smem_u32(smem), 2048, 128), db = tc_desc(tc_smem_u32(smem) + (4 << 10), N * 16, 128);
    for (int i = 0; i < iters; i += 2) {
      tc_mma(tmem, da, db, idesc, i > 0);
      tc_mma(tmem + 256, da, db, idesc, i > 0);
    }
    tc_commit(&bar);
    tc_mbar_wait(&bar, 0, 700);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace esm

using namespace esm;

extern "C" int esm_umma_tf32_peak(int iters, float* tflops, void* stream) {
  ESM_REQUIRE(tflops && iters >= 2, "umma_tf32_peak: bad arguments");
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    set_error("umma_tf32_peak: no CUDA device");
    return ESM_ERR_CUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  iters &= ~1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 0.f;
  for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms
    cudaEventRecord(e0, st);
    umma_tf32_peak_kernel<<<sms, 160, 16 << 10, st>>>(iters);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      return check_launch("umma_tf32_peak");
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const float tf = (float)((double)sms * iters * 2.0 * 128 * 256 * 8 / (ms * 1e-3) / 1e12);
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return check_launch("umma_tf32_peak");
}
