// Kernels the timm-compatible backbones (esmstereo_b200/timm_compat.py: EfficientNet-B2, MobileNetV2, the `Feature`
// taps of ESMStereo.py:40-77) need besides the conv engines: depthwise k x k convolution + folded BN + activation,
// global average pooling and the per-(image, channel) gate of squeeze-and-excitation.  All HBM-bound, NCHW fp32.
#include "common.cuh"

namespace esm {

// y[b, c, yo, xo] = act(scale[c] * sum_{i,j} x[b, c, yo*s + i - p, xo*s + j - p] * w[c, i, j] + shift[c]),  p = K / 2
template <int K>
__global__ void __launch_bounds__(256) dwconv2d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, int act, float* __restrict__ y, int C, int H, int W, int Ho,
                                                       int Wo, int stride, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xo = (int)(i % Wo);
  long long t = i / Wo;
  const int yo = (int)(t % Ho);
  t /= Ho;
  const int c = (int)(t % C);
  const float* xp = x + t * (long long)H * W;  // plane (b, c)
  const float* wp = w + (long long)c * K * K;
  constexpr int P = K / 2;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    const int yi = yo * stride + a - P;
    if ((unsigned)yi >= (unsigned)H) continue;
#pragma unroll
    for (int b2 = 0; b2 < K; ++b2) {
      const int xi = xo * stride + b2 - P;
      if ((unsigned)xi < (unsigned)W) acc = fmaf(__ldg(xp + (long long)yi * W + xi), __ldg(wp + a * K + b2), acc);
    }
  }
  const float sc = scale ? __ldg(scale + c) : 1.f, sh = shift ? __ldg(shift + c) : 0.f;
  y[i] = apply_act(fmaf(acc, sc, sh), act);
}

__global__ void __launch_bounds__(256) global_avgpool_kernel(const float* __restrict__ x, float* __restrict__ out, int HW) {
  __shared__ float part[8];
  const float* p = x + (long long)blockIdx.x * HW;
  float s = 0.f;
  for (int i = threadIdx.x; i < HW; i += 256) s += __ldg(p + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    out[blockIdx.x] = t / (float)HW;
  }
}

__global__ void __launch_bounds__(256) scale_channels_kernel(float* __restrict__ x, const float* __restrict__ gate, int HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) x[i] *= __ldg(gate + i / HW);
}

}  // namespace esm

using namespace esm;

extern "C" int esm_dwconv2d_f32(const float* x, const float* w, const float* scale, const float* shift, int act, float* y, int B, int C, int H,
                                int W, int k, int stride, void* stream) {
  ESM_REQUIRE(x && w && y, "dwconv2d: null pointer");
  ESM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "dwconv2d: empty shape");
  ESM_REQUIRE((k == 3 || k == 5 || k == 7) && (stride == 1 || stride == 2), "dwconv2d: k must be 3, 5 or 7 and stride 1 or 2");
  const int p = k / 2, Ho = (H + 2 * p - k) / stride + 1, Wo = (W + 2 * p - k) / stride + 1;
  const long long total = (long long)B * C * Ho * Wo;
  ESM_REQUIRE(ceil_div_ll(total, 256) < (1ll << 31), "dwconv2d: grid too large");
  const unsigned grid = (unsigned)ceil_div_ll(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 3)
    dwconv2d_kernel<3><<<grid, 256, 0, st>>>(x, w, scale, shift, act, y, C, H, W, Ho, Wo, stride, total);
  else if (k == 5)
    dwconv2d_kernel<5><<<grid, 256, 0, st>>>(x, w, scale, shift, act, y, C, H, W, Ho, Wo, stride, total);
  else
    dwconv2d_kernel<7><<<grid, 256, 0, st>>>(x, w, scale, shift, act, y, C, H, W, Ho, Wo, stride, total);
  return check_launch("dwconv2d");
}

extern "C" int esm_global_avgpool_f32(const float* x, float* out, int B, int C, int HW, void* stream) {
  ESM_REQUIRE(x && out && B > 0 && C > 0 && HW > 0, "global_avgpool: bad arguments");
  global_avgpool_kernel<<<(unsigned)(B * C), 256, 0, (cudaStream_t)stream>>>(x, out, HW);
  return check_launch("global_avgpool");
}

extern "C" int esm_scale_channels_f32(float* x, const float* gate, int B, int C, int HW, void* stream) {
  ESM_REQUIRE(x && gate && B > 0 && C > 0 && HW > 0, "scale_channels: bad arguments");
  const long long total = (long long)B * C * HW;
  scale_channels_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(x, gate, HW, total);
  return check_launch("scale_channels");
}
