// Error plumbing and small utilities of the C ABI (include/esm_b200.h).
#include "common.cuh"

#include <stdarg.h>

namespace esm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the (non-sticky) launch error so the next call starts clean
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ESM_ERR_CUDA;
  }
  return ESM_OK;
}

__global__ void fill_kernel(float* p, long long n, float v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace esm

using namespace esm;

extern "C" const char* esm_last_error(void) { return g_err; }
extern "C" int esm_version(void) { return 100; }

extern "C" int esm_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("device_info: no CUDA device");
    return ESM_ERR_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("device_info: cudaGetDeviceProperties failed");
    return ESM_ERR_CUDA;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return ESM_OK;
}

extern "C" int esm_fill_f32(float* p, long long n, float value, void* stream) {
  ESM_REQUIRE(p && n > 0, "fill: null pointer or empty range");
  fill_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(p, n, value);
  return check_launch("fill");
}

// Synchronous device -> host copy of raw bytes (esmstereo_b200/engine.py snapshots the weights of an engine with it).
extern "C" int esm_download(void* dst_host, const void* src_device, long long nbytes) {
  ESM_REQUIRE(dst_host && src_device && nbytes > 0, "download: bad arguments");
  if (cudaMemcpy(dst_host, src_device, (size_t)nbytes, cudaMemcpyDeviceToHost) != cudaSuccess) return check_launch("download");
  return ESM_OK;
}
