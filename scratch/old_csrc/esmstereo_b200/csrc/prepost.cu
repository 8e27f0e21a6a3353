// Image pre-processing and disparity post-processing on the device (the per-image CPU work
// either side of the forward in the reference's scripts).
//   esm_preprocess_u8_f32   <- transforms.ToTensor + Normalize(ImageNet) + pad to the network size
//                              (test_kitti.py:93-106: PIL crop with negative origin = black pixels on the top / left,
//                               normalised like any pixel; datasets/kitti_dataset.py:145-160: zeros AFTER normalisation
//                               on the top / right)
//   esm_postprocess_disp_u16 <- un-pad + round(d * 256) -> uint16 (test_kitti.py:114,127; save_disp.py:83-88)
// Both are one pass over the image, bound by HBM (3 B read + 12 B written per pixel; 4 B read + 2 B written).
#include "common.cuh"

namespace esm {

// One thread per output pixel (all 3 channels): the HWC bytes of a pixel are read once, the three CHW planes are
// written with coalesced stores.
__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ rgb, float* __restrict__ out, int h, int w, int Hp, int Wp,
                                                         int pad_top, int pad_left, int fill_normalised, float m0, float m1, float m2, float s0,
                                                         float s1, float s2, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % Wp);
  const long long t = i / Wp;
  const int y = (int)(t % Hp);
  const long long b = t / Hp;
  const int sy = y - pad_top, sx = x - pad_left;
  const bool in = (unsigned)sy < (unsigned)h && (unsigned)sx < (unsigned)w;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (in) {
    const uint8_t* px = rgb + ((b * h + sy) * (long long)w + sx) * 3;
    v0 = (float)px[0];
    v1 = (float)px[1];
    v2 = (float)px[2];
  }
  float* o = out + (b * 3 * Hp + y) * (long long)Wp + x;
  const long long plane = (long long)Hp * Wp;
  if (in || fill_normalised) {
    // ToTensor: uint8 -> float32 / 255; Normalize: (t - mean) / std -- true divisions, like torch
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(v0, 255.0f), m0), s0);
    o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(v1, 255.0f), m1), s1);
    o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(v2, 255.0f), m2), s2);
  } else {
    o[0] = 0.f;
    o[plane] = 0.f;
    o[2 * plane] = 0.f;
  }
}

__global__ void __launch_bounds__(256) postprocess_kernel(const float* __restrict__ disp, uint16_t* __restrict__ out, int Hp, int Wp, int top,
                                                          int left, int h, int w, float scale, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % w);
  const long long t = i / w;
  const int y = (int)(t % h);
  const long long b = t / h;
  const float d = __ldg(disp + (b * Hp + top + y) * (long long)Wp + left + x);
  // np.round = round half to even; the cast saturates instead of wrapping (disparities are non-negative)
  const float r = rintf(__fmul_rn(d, scale));
  out[i] = (uint16_t)fminf(fmaxf(r, 0.f), 65535.f);
}

// [B, 8 = (pd, ph, pw), D2, H2, W2] (strided) -> [B, 1, 2 D2, 2 H2, 2 W2] contiguous: the PixelShuffle of the sub-pixel
// form of a one-channel ConvTranspose3d k4 s2 p1 (ESMStereo.py:150: `conv1_up` of the hourglass)
__global__ void __launch_bounds__(256) pixel_shuffle3d_kernel(const float* __restrict__ in, long long sB, long long sC, long long sD, long long sH,
                                                              float* __restrict__ out, int D2, int H2, int W2, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W = 2 * W2, H = 2 * H2, D = 2 * D2;
  const int x = (int)(i % W);
  long long t = i / W;
  const int y = (int)(t % H);
  t /= H;
  const int d = (int)(t % D);
  const long long b = t / D;
  out[i] = __ldg(in + b * sB + (long long)((d & 1) * 4 + (y & 1) * 2 + (x & 1)) * sC + (long long)(d >> 1) * sD + (long long)(y >> 1) * sH + (x >> 1));
}

// Post-processing of the reference's ROS publisher (kitti_publisher/src/kitti_publisher_cuda_node.cpp:385-404): crop the
// padded disparity to the image, cv::medianBlur(5) (float: exact median of the 5 x 5 window, replicated border), zero
// where not 0 < d < max_disp, convertTo(CV_16UC1, scale) = saturate(round-half-even(d * scale)).
__global__ void __launch_bounds__(256) publish_kernel(const float* __restrict__ disp, uint16_t* __restrict__ out, int Wp, int h, int w,
                                                      float max_disp, float scale, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % w), y = (int)(i / w);
  float v[25];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    const int yy = min(max(y + a - 2, 0), h - 1);
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      const int xx = min(max(x + b - 2, 0), w - 1);
      v[a * 5 + b] = __ldg(disp + (long long)yy * Wp + xx);
    }
  }
  // partial selection sort up to the 13th smallest
#pragma unroll
  for (int a = 0; a < 13; ++a) {
#pragma unroll
    for (int b = a + 1; b < 25; ++b) {
      const float lo = fminf(v[a], v[b]), hi = fmaxf(v[a], v[b]);
      v[a] = lo;
      v[b] = hi;
    }
  }
  float m = v[12];
  if (!(m > 0.f && m < max_disp)) m = 0.f;
  out[i] = (uint16_t)fminf(fmaxf(rintf(__fmul_rn(m, scale)), 0.f), 65535.f);
}

}  // namespace esm

using namespace esm;

extern "C" int esm_copy_f32(float* dst, const float* src, long long n, void* stream) {
  ESM_REQUIRE(dst && src && n > 0, "copy: null pointer or empty range");
  if (cudaMemcpyAsync(dst, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) return check_launch("copy");
  return ESM_OK;
}

extern "C" int esm_pixel_shuffle3d_f32(const float* in, long long sB, long long sC, long long sD, long long sH, float* out, int B, int D2,
                                       int H2, int W2, void* stream) {
  ESM_REQUIRE(in && out && B > 0 && D2 > 0 && H2 > 0 && W2 > 0, "pixel_shuffle3d: bad arguments");
  const long long total = 8ll * B * D2 * H2 * W2;
  pixel_shuffle3d_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(in, sB, sC, sD, sH, out, D2, H2, W2, total);
  return check_launch("pixel_shuffle3d");
}

extern "C" int esm_disparity_publish_u16(const float* disp, unsigned short* out, int Hp, int Wp, int h, int w, float max_disp, float scale,
                                         void* stream) {
  ESM_REQUIRE(disp && out && h > 0 && w > 0 && h <= Hp && w <= Wp, "disparity_publish: bad arguments");
  const long long total = (long long)h * w;
  publish_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(disp, out, Wp, h, w, max_disp, scale, total);
  return check_launch("disparity_publish");
}

extern "C" int esm_preprocess_u8_f32(const unsigned char* rgb_hwc, float* out_chw, int B, int h, int w, int Hp, int Wp, int pad_top,
                                     int pad_left, int fill_normalised, const float* mean3, const float* std3, void* stream) {
  ESM_REQUIRE(rgb_hwc && out_chw && mean3 && std3, "preprocess: null pointer");
  ESM_REQUIRE(B > 0 && h > 0 && w > 0 && Hp >= h && Wp >= w, "preprocess: bad shape (%d x %d into %d x %d)", h, w, Hp, Wp);
  ESM_REQUIRE(pad_top >= 0 && pad_left >= 0 && pad_top + h <= Hp && pad_left + w <= Wp, "preprocess: the image does not fit at (%d, %d)",
              pad_top, pad_left);
  const long long total = (long long)B * Hp * Wp;
  preprocess_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(rgb_hwc, out_chw, h, w, Hp, Wp, pad_top, pad_left,
                                                                                         fill_normalised, mean3[0], mean3[1], mean3[2], std3[0],
                                                                                         std3[1], std3[2], total);
  return check_launch("preprocess");
}

extern "C" int esm_postprocess_disp_u16(const float* disp, unsigned short* out, int B, int Hp, int Wp, int top, int left, int h, int w,
                                        float scale, void* stream) {
  ESM_REQUIRE(disp && out, "postprocess: null pointer");
  ESM_REQUIRE(B > 0 && h > 0 && w > 0 && top >= 0 && left >= 0 && top + h <= Hp && left + w <= Wp, "postprocess: crop outside the disparity map");
  const long long total = (long long)B * h * w;
  postprocess_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(disp, out, Hp, Wp, top, left, h, w, scale, total);
  return check_launch("postprocess");
}
