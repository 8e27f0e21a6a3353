// Disparity regression and final assembly (bandwidth-bound, one pass over the cost volume).
//   esm_regression_top2_f32       <- regression_topk(k=2) (submodule.py:218-225; ESMStereo.py:719-721)
//   esm_disparity_regression_f32  <- disparity_regression (submodule.py:211-216)  [no softmax]
//   esm_bilinear_add_f32          <- F.interpolate(bilinear, align_corners=False) + residual, * scale
//                                    (ESMStereo.py:307,316,497,507,745)
// The reference sorts all D costs per pixel, gathers twice and materialises a `disp_samples`
// tensor; here each thread streams its pixel's D costs once, keeping a running top-2 in registers.
#include "common.cuh"

namespace esm {

// One CTA = 32 consecutive pixels x 4 disparity ranges (one warp each): a thread has its whole range (<= 16 values)
// in flight at once, so a pixel costs one DRAM round trip instead of D/16 (the first version: one thread per pixel, 16
// loads at a time -- 8 us for 5.9 MB at KITTI shape, all of it latency); the four partial top-2 lists are merged
// through shared memory.  Descending stable order: strict '>' keeps the lower index first on ties.
struct Top2 {
  float v1, v2;
  int i1, i2;  // -1: empty slot
};
__device__ __forceinline__ void top2_push(Top2& t, float x, int i) {
  if (t.i1 < 0 || x > t.v1) {
    t.v2 = t.v1; t.i2 = t.i1;
    t.v1 = x; t.i1 = i;
  } else if (t.i2 < 0 || x > t.v2) {
    t.v2 = x; t.i2 = i;
  }
}
constexpr int REG_SPLIT = 4, REG_MAXR = 16;

// SUB: the cost volume is read in the sub-pixel form the hourglass's last layer produces it in -- `conv1_up`
// (ConvTranspose3d k4 s2 p1 to one channel, ESMStereo.py:150,182) runs as a k3 convolution to 8 phase channels
// [B, 8 = (pd, ph, pw), D/2, H/2, W/2] -- so the PixelShuffle copy of the volume never happens:
//   cost[b, d, y, x] = y8[b, (d&1)*4 + (y&1)*2 + (x&1), d>>1, y>>1, x>>1]
struct SubSrc {
  long long sB, sC, sD, sH;  // strides of y8 in elements (W stride 1)
  int W;                     // full-resolution width (2 * W2)
};
template <bool SUB>
__global__ void __launch_bounds__(128) regression_top2_kernel(const float* __restrict__ cost, float* __restrict__ pred,
                                                              int* __restrict__ idx, int D, long long plane,
                                                              long long total, SubSrc ss) {
  __shared__ Top2 s_part[REG_SPLIT - 1][32];
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  const bool live = i < total;
  const long long b = live ? i / plane : 0;
  const long long p = live ? i - b * plane : 0;
  const float* c = cost + b * D * plane + p;
  long long sub_d = 0, sub_odd = 0;  // SUB: address of disparity d = c + (d >> 1) * sub_d + (d & 1) * sub_odd
  if (SUB) {
    const int y = (int)(p / ss.W), x = (int)(p - (long long)y * ss.W);
    c = cost + b * ss.sB + (long long)((y & 1) * 2 + (x & 1)) * ss.sC + (long long)(y >> 1) * ss.sH + (x >> 1);
    sub_d = ss.sD;
    sub_odd = 4 * ss.sC;
  }
  const int per = (D + REG_SPLIT - 1) / REG_SPLIT;  // disparities per warp
  Top2 t = {-INFINITY, -INFINITY, -1, -1};
  for (int d0 = part * per; d0 < min(D, (part + 1) * per); d0 += REG_MAXR) {
    const int dend = min(D, (part + 1) * per);
    float v[REG_MAXR];
#pragma unroll
    for (int u = 0; u < REG_MAXR; ++u) {
      const int d = d0 + u;
      const long long off = SUB ? (long long)(d >> 1) * sub_d + (long long)(d & 1) * sub_odd : (long long)d * plane;
      v[u] = (live && d < dend) ? __ldg(c + off) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < REG_MAXR; ++u)
      if (d0 + u < dend) top2_push(t, v[u], d0 + u);
  }
  if (part > 0) s_part[part - 1][lane] = t;
  __syncthreads();
  if (part > 0 || !live) return;
#pragma unroll
  for (int q = 0; q < REG_SPLIT - 1; ++q) {  // ranges in ascending disparity order: earlier entries win ties
    const Top2 o = s_part[q][lane];
    if (o.i1 >= 0) top2_push(t, o.v1, o.i1);
    if (o.i2 >= 0) top2_push(t, o.v2, o.i2);
  }
  const float v1 = t.v1, v2 = t.v2;
  const int i1 = t.i1, i2 = t.i2;
  float out;
  if (D >= 2) {
    // softmax over (v1, v2) as torch does it: exp(x - max) / sum
    const float e2 = expf(v2 - v1);
    const float s = 1.0f + e2;
    const float p1 = 1.0f / s, p2 = e2 / s;
    out = __fadd_rn(__fmul_rn((float)i1, p1), __fmul_rn((float)i2, p2));
  } else {
    out = (float)i1;
  }
  pred[i] = out;
  if (idx) {
    idx[(b * 2 + 0) * plane + p] = i1;
    idx[(b * 2 + 1) * plane + p] = (D >= 2) ? i2 : i1;
  }
}

__global__ void __launch_bounds__(128) disparity_regression_kernel(const float* __restrict__ cost,
                                                                   float* __restrict__ pred, int D, long long plane,
                                                                   long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / plane;
  const long long p = i - b * plane;
  const float* c = cost + b * D * plane + p;
  float s = 0.f;
  constexpr int U = 16;
  for (int d0 = 0; d0 < D; d0 += U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = (d0 + u < D) ? __ldg(c + (long long)(d0 + u) * plane) : 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (d0 + u < D) s = __fadd_rn(s, __fmul_rn(v[u], (float)(d0 + u)));
  }
  pred[i] = s;
}

// torch upsample_bilinear2d, align_corners=False, scale_factor = factor (src = (dst+0.5)/f - 0.5, clamped at 0)
__global__ void __launch_bounds__(256) bilinear_add_kernel(const float* __restrict__ prev, const float* __restrict__ res,
                                                           float* __restrict__ out, int h, int w, int f, float rscale,
                                                           float out_scale, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W = w * f, H = h * f;
  const int x = (int)(i % W);
  const long long t = i / W;
  const int y = (int)(t % H);
  const long long b = t / H;
  float sy = ((float)y + 0.5f) * rscale - 0.5f;
  float sx = ((float)x + 0.5f) * rscale - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + ((y0 < h - 1) ? 1 : 0), x1 = x0 + ((x0 < w - 1) ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float* pb = prev + b * (long long)h * w;
  const float v00 = __ldg(pb + (long long)y0 * w + x0), v01 = __ldg(pb + (long long)y0 * w + x1);
  const float v10 = __ldg(pb + (long long)y1 * w + x0), v11 = __ldg(pb + (long long)y1 * w + x1);
  const float up = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  out[i] = (up + __ldg(res + i)) * out_scale;
}

}  // namespace esm

using namespace esm;

extern "C" int esm_regression_top2_f32(const float* cost, float* pred, int* idx, int B, int D, int H, int W,
                                       void* stream) {
  ESM_REQUIRE(cost && pred, "regression_top2: null pointer");
  ESM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "regression_top2: empty shape");
  const long long plane = (long long)H * W, total = plane * B;
  regression_top2_kernel<false><<<(unsigned)ceil_div_ll(total, 32), 128, 0, (cudaStream_t)stream>>>(cost, pred, idx, D, plane, total, SubSrc{});
  return check_launch("regression_top2");
}

extern "C" int esm_regression_top2_subpixel_f32(const float* y8, long long sB, long long sC, long long sD, long long sH, float* pred, int* idx,
                                                int B, int D2, int H2, int W2, void* stream) {
  ESM_REQUIRE(y8 && pred, "regression_top2_subpixel: null pointer");
  ESM_REQUIRE(B > 0 && D2 > 0 && H2 > 0 && W2 > 0, "regression_top2_subpixel: empty shape");
  const long long plane = 4ll * H2 * W2, total = plane * B;
  SubSrc ss = {sB, sC, sD, sH, 2 * W2};
  regression_top2_kernel<true><<<(unsigned)ceil_div_ll(total, 32), 128, 0, (cudaStream_t)stream>>>(y8, pred, idx, 2 * D2, plane, total, ss);
  return check_launch("regression_top2_subpixel");
}

extern "C" int esm_disparity_regression_f32(const float* cost, float* pred, int B, int D, int H, int W, void* stream) {
  ESM_REQUIRE(cost && pred, "disparity_regression: null pointer");
  ESM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "disparity_regression: empty shape");
  const long long plane = (long long)H * W, total = plane * B;
  disparity_regression_kernel<<<(unsigned)ceil_div_ll(total, 128), 128, 0, (cudaStream_t)stream>>>(cost, pred, D, plane,
                                                                                                   total);
  return check_launch("disparity_regression");
}

extern "C" int esm_bilinear_add_f32(const float* prev, const float* residual, float* out, int B, int h, int w,
                                    int factor, float out_scale, void* stream) {
  ESM_REQUIRE(prev && residual && out, "bilinear_add: null pointer");
  ESM_REQUIRE(B > 0 && h > 0 && w > 0 && factor >= 1, "bilinear_add: empty shape");
  const long long total = (long long)B * h * factor * w * factor;
  bilinear_add_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
      prev, residual, out, h, w, factor, 1.0f / (float)factor, out_scale, total);
  return check_launch("bilinear_add");
}
