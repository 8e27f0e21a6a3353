// Confidence head (LAFNet_ESM / conf_upsample, ESMStereo_confidence.py:511-744): the pieces that are not
// plain convolutions, each fused into one bandwidth-bound kernel.
#include "common.cuh"

namespace esm {

// softmax(-cost/||cost|| * 100) over D, 7 largest probabilities descending (ESMStereo_confidence.py:645-653)
template <int MAXD>
__global__ void __launch_bounds__(128) laf_cost_top7_kernel(const float* __restrict__ cost, float* __restrict__ out,
                                                            int D, long long plane, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / plane;
  const long long p = i - b * plane;
  const float* c = cost + b * D * plane + p;
  float v[MAXD];
  float ss = 0.f;
#pragma unroll
  for (int d = 0; d < MAXD; ++d) {
    v[d] = (d < D) ? __ldg(c + (long long)d * plane) : 0.f;
    ss = __fadd_rn(ss, __fmul_rn(v[d], v[d]));
  }
  const float nrm = sqrtf(ss + 1e-6f);
  float mx = -INFINITY;
#pragma unroll
  for (int d = 0; d < MAXD; ++d) {
    v[d] = (d < D) ? -(v[d] / nrm) * 100.0f : -INFINITY;
    mx = fmaxf(mx, v[d]);
  }
  float sum = 0.f;
#pragma unroll
  for (int d = 0; d < MAXD; ++d) {
    v[d] = (d < D) ? expf(v[d] - mx) : -1.0f;  // -1 marks padding (never selected before real entries)
    if (d < D) sum += v[d];
  }
  // selection of the 7 largest (values only)
  float* o = out + b * 7 * plane + p;
  for (int k = 0; k < 7; ++k) {
    float best = -2.0f;
    int bi = 0;
#pragma unroll
    for (int d = 0; d < MAXD; ++d)
      if (v[d] > best) {
        best = v[d];
        bi = d;
      }
#pragma unroll
    for (int d = 0; d < MAXD; ++d)
      if (d == bi) v[d] = -3.0f;
    o[(long long)k * plane] = best / sum;
  }
}

// softmax over the three 1-channel attention maps; towers scaled and concatenated (":676-686")
__global__ void __launch_bounds__(256) laf_attention_kernel(const float* __restrict__ cx, const float* __restrict__ dx,
                                                            const float* __restrict__ ix, const float* __restrict__ ac,
                                                            const float* __restrict__ ad, const float* __restrict__ ai,
                                                            float* __restrict__ out, int C, long long plane,
                                                            long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / plane;
  const long long p = i - b * plane;
  const float a0 = __ldg(ac + i), a1 = __ldg(ad + i), a2 = __ldg(ai + i);
  const float mx = fmaxf(a0, fmaxf(a1, a2));
  const float e0 = expf(a0 - mx), e1 = expf(a1 - mx), e2 = expf(a2 - mx);
  const float s = e0 + e1 + e2;
  const float w0 = e0 / s, w1 = e1 / s, w2 = e2 / s;
  const long long ib = b * C * plane + p;
  float* o = out + b * 3 * C * plane + p;
  for (int c = 0; c < C; ++c) {
    o[(long long)c * plane] = __ldg(cx + ib + c * plane) * w0;
    o[(long long)(C + c) * plane] = __ldg(dx + ib + c * plane) * w1;
    o[(long long)(2 * C + c) * plane] = __ldg(ix + ib + c * plane) * w2;
  }
}

// bilinear grid_sample (align_corners=True, zero padding) of one channel plane
__device__ __forceinline__ float sample_zero(const float* __restrict__ f, int H, int W, float px, float py) {
  const float fx = floorf(px), fy = floorf(py);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = px - fx, wy1 = py - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;  // = (ix_se - ix) etc.
  float r = 0.f;
  const bool xin0 = x0 >= 0 && x0 < W, xin1 = x1 >= 0 && x1 < W;
  const bool yin0 = y0 >= 0 && y0 < H, yin1 = y1 >= 0 && y1 < H;
  if (yin0 && xin0) r += __ldg(f + (long long)y0 * W + x0) * (wx0 * wy0);
  if (yin0 && xin1) r += __ldg(f + (long long)y0 * W + x1) * (wx1 * wy0);
  if (yin1 && xin0) r += __ldg(f + (long long)y1 * W + x0) * (wx0 * wy1);
  if (yin1 && xin1) r += __ldg(f + (long long)y1 * W + x1) * (wx1 * wy1);
  return r;
}

// scale-adaptive 3x3 sampling + embed_conv2 (k3 s3) + BN + ReLU, ESMStereo_confidence.py:693-719
template <int C>
__global__ void __launch_bounds__(128) laf_sample_embed_kernel(const float* __restrict__ feat,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ lin_x,
                                                               const float* __restrict__ lin_y,
                                                               const float* __restrict__ weight,
                                                               const float* __restrict__ bn_scale,
                                                               const float* __restrict__ bn_shift,
                                                               float* __restrict__ out, int H, int W, float step_y,
                                                               long long total) {
  __shared__ float wsm[C * C * 9];  // [co][ci][ky][kx] (torch layout)
  for (int i = threadIdx.x; i < C * C * 9; i += blockDim.x) wsm[i] = weight[i];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long plane = (long long)H * W;
  const long long b = i / plane;
  const long long p = i - b * plane;
  const int y = (int)(p / W), x = (int)(p - (long long)y * W);
  const float sc = __ldg(scale + i);
  const float gx = __ldg(lin_x + x), gy = __ldg(lin_y + y);
  const float* fb = feat + b * C * plane;
  float acc[C];
#pragma unroll
  for (int co = 0; co < C; ++co) acc[co] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int oy = t / 3 - 1, ox = t % 3 - 1;
    // grid + cat((ox*step_y*scale, oy*scale)): the python scalar is rounded to fp32 before the multiply
    const float nx = gx + ((float)ox * step_y) * sc;
    const float ny = gy + ((float)oy) * sc;
    const float px = ((nx + 1.f) / 2.f) * (float)(W - 1);  // unnormalize, align_corners=True
    const float py = ((ny + 1.f) / 2.f) * (float)(H - 1);
    for (int ci = 0; ci < C; ++ci) {
      const float v = sample_zero(fb + ci * plane, H, W, px, py);
#pragma unroll
      for (int co = 0; co < C; ++co) acc[co] = fmaf(wsm[(co * C + ci) * 9 + t], v, acc[co]);
    }
  }
  float* o = out + b * C * plane + p;
#pragma unroll
  for (int co = 0; co < C; ++co)
    o[(long long)co * plane] = fmaxf(fmaf(acc[co], __ldg(bn_scale + co), __ldg(bn_shift + co)), 0.f);
}

// ConvTranspose2d(C->9,k4,s4) + softmax(9) + convex combination of the 3x3 neighbourhood (":536-543")
__global__ void __launch_bounds__(256) conf_convex_up4_kernel(const float* __restrict__ feat,
                                                              const float* __restrict__ conf,
                                                              const float* __restrict__ weight,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int C, int h, int w, long long total) {
  extern __shared__ float wsm[];  // [C][9][4][4] + bias[9]
  for (int i = threadIdx.x; i < C * 144; i += blockDim.x) wsm[i] = weight[i];
  if (threadIdx.x < 9) wsm[C * 144 + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W4 = w * 4, H4 = h * 4;
  const int X = (int)(i % W4);
  const long long t = i / W4;
  const int Y = (int)(t % H4);
  const long long b = t / H4;
  const int yi = Y >> 2, a = Y & 3, xj = X >> 2, bb = X & 3;
  const long long plane = (long long)h * w;
  const float* fb = feat + b * C * plane + (long long)yi * w + xj;
  float lg[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) lg[k] = 0.f;
  for (int c = 0; c < C; ++c) {
    const float f = __ldg(fb + c * plane);
#pragma unroll
    for (int k = 0; k < 9; ++k) lg[k] = fmaf(f, wsm[((c * 9 + k) * 4 + a) * 4 + bb], lg[k]);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    lg[k] += wsm[C * 144 + k];
    mx = fmaxf(mx, lg[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    lg[k] = expf(lg[k] - mx);
    s += lg[k];
  }
  const float* cb = conf + b * plane;
  float r = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = yi + k / 3 - 1, xx = xj + k % 3 - 1;
    const float cv = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(cb + (long long)yy * w + xx) : 0.f;
    r += cv * (lg[k] / s);
  }
  out[i] = r;
}

}  // namespace esm

using namespace esm;

extern "C" int esm_laf_cost_top7_f32(const float* cost, float* out, int B, int D, int H, int W, void* stream) {
  ESM_REQUIRE(cost && out && B > 0 && H > 0 && W > 0, "laf_cost_top7: null pointer or empty shape");
  ESM_REQUIRE(D >= 7 && D <= 64, "laf_cost_top7: D must be in [7,64] (got %d)", D);
  const long long plane = (long long)H * W, total = plane * B;
  const unsigned grid = (unsigned)ceil_div_ll(total, 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (D <= 12)
    laf_cost_top7_kernel<12><<<grid, 128, 0, st>>>(cost, out, D, plane, total);
  else if (D <= 24)
    laf_cost_top7_kernel<24><<<grid, 128, 0, st>>>(cost, out, D, plane, total);
  else
    laf_cost_top7_kernel<64><<<grid, 128, 0, st>>>(cost, out, D, plane, total);
  return check_launch("laf_cost_top7");
}

extern "C" int esm_laf_attention_f32(const float* cost_x, const float* disp_x, const float* imag_x, const float* att_c,
                                     const float* att_d, const float* att_i, float* out, int B, int C, int H, int W,
                                     void* stream) {
  ESM_REQUIRE(cost_x && disp_x && imag_x && att_c && att_d && att_i && out, "laf_attention: null pointer");
  ESM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "laf_attention: empty shape");
  const long long plane = (long long)H * W, total = plane * B;
  laf_attention_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
      cost_x, disp_x, imag_x, att_c, att_d, att_i, out, C, plane, total);
  return check_launch("laf_attention");
}

extern "C" int esm_laf_sample_embed_f32(const float* feat, const float* scale, const float* lin_x, const float* lin_y,
                                        const float* weight, const float* bn_scale, const float* bn_shift, float* out,
                                        int B, int C, int H, int W, void* stream) {
  ESM_REQUIRE(feat && scale && lin_x && lin_y && weight && bn_scale && bn_shift && out, "laf_sample_embed: null pointer");
  ESM_REQUIRE(B > 0 && H > 0 && W > 1, "laf_sample_embed: empty shape (W must be > 1)");
  ESM_REQUIRE(C == 16, "laf_sample_embed: C must be 16 (LAFNet_ESM(16), ESMStereo_confidence.py:871)");
  const long long total = (long long)B * H * W;
  const float step_y = (float)(2.0 / (double)(W - 1));  // ":705"; step_x (":704") is unused by the reference
  laf_sample_embed_kernel<16><<<(unsigned)ceil_div_ll(total, 128), 128, 0, (cudaStream_t)stream>>>(
      feat, scale, lin_x, lin_y, weight, bn_scale, bn_shift, out, H, W, step_y, total);
  return check_launch("laf_sample_embed");
}

extern "C" int esm_conf_convex_up4_f32(const float* feat, const float* conf, const float* weight, const float* bias,
                                       float* out, int B, int C, int h, int w, void* stream) {
  ESM_REQUIRE(feat && conf && weight && bias && out, "conf_convex_up4: null pointer");
  ESM_REQUIRE(B > 0 && C > 0 && C <= 64 && h > 0 && w > 0, "conf_convex_up4: bad shape");
  const long long total = (long long)B * h * 4 * w * 4;
  const size_t smem = ((size_t)C * 144 + 9) * sizeof(float);
  conf_convex_up4_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, smem, (cudaStream_t)stream>>>(feat, conf, weight, bias,
                                                                                                 out, C, h, w, total);
  return check_launch("conf_convex_up4");
}
