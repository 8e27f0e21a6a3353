// Cost-volume builders (HBM-write-bound streaming kernels).
//   esm_gwc_volume_f32        <- build_gwc_volume + groupwise_correlation (submodule.py:143-161)
//   esm_norm_corr_volume_f32  <- build_norm_correlation_volume + norm_correlation (submodule.py:187-200)
//   esm_concat_volume_f32     <- build_concat_volume (submodule.py:129-140)
//   esm_substract_volume_f32  <- build_substract_volume + groupwise_difference (submodule.py:104-126)
//   esm_gwc_volume_norm_f32   <- build_gwc_volume_norm + groupwise_correlation_norm (submodule.py:163-184)
//   (the last two are not used by any model configuration)
// The reference builds the volume with a Python loop over disparities (memset + D x {mul, mean,
// strided copy}); here one launch writes every output element exactly once, zeros included.
//
// Algorithmic bytes: 4*(2*C*h*w + G*D*h*w) per pair; the kernel is bound by the
// G*D*h*w fp32 stores.  Thread layout: a thread owns 4 consecutive x of one row and GPT groups,
// keeps its left-feature values in registers and slides a register window over the right row as d
// grows (one new right value per channel per disparity), so every store is a coalesced STG.128 and
// L/R are read from L1/L2 once per thread.
#include "common.cuh"

namespace esm {

// GPT = groups per thread, CPG = channels per group (template for register residency)
template <int GPT, int CPG>
__global__ void __launch_bounds__(256) gwc_volume_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                                         float* __restrict__ V, int C, int H, int W, int D, int G,
                                                         int xg_per_row, long long positions) {
  const int lane_pos = blockIdx.x * 32 + (threadIdx.x & 31);  // (y, x-group) position
  const int gset = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  const int g0 = gset * GPT;
  if (lane_pos >= positions || g0 >= G) return;
  const int y = lane_pos / xg_per_row;
  const int x0 = (lane_pos - y * xg_per_row) * 4;
  const long long plane = (long long)H * W;
  const float* Lb = L + ((long long)b * C + (long long)g0 * CPG) * plane + (long long)y * W;
  const float* Rb = R + ((long long)b * C + (long long)g0 * CPG) * plane + (long long)y * W;

  float l[GPT * CPG][4];
  float r[GPT * CPG][4];  // r[.][i] = R[x0 + i - d]
#pragma unroll
  for (int c = 0; c < GPT * CPG; ++c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = x0 + i;
      const bool in = x < W;
      l[c][i] = in ? __ldg(Lb + c * plane + x) : 0.f;
      r[c][i] = in ? __ldg(Rb + c * plane + x) : 0.f;
    }
  }
  const bool vec = ((W & 3) == 0) && (x0 + 4 <= W);
  float* Vb = V + (((long long)b * G + g0) * D) * plane + (long long)y * W + x0;
  const long long gstride = (long long)D * plane;
  for (int d = 0; d < D; ++d) {
#pragma unroll
    for (int g = 0; g < GPT; ++g) {
      if (g0 + g >= G) break;
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < CPG; ++k) s = __fadd_rn(s, __fmul_rn(l[g * CPG + k][i], r[g * CPG + k][i]));
        o[i] = (x0 + i >= d) ? s / (float)CPG : 0.f;  // zero triangle x < d, submodule.py:153,156
      }
      float* dst = Vb + g * gstride + (long long)d * plane;
      if (vec) {
        __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x0 + i < W) dst[i] = o[i];
      }
    }
    // slide the right-image window by one column: r[.][i] <- R[x0 + i - (d+1)]
    const int xn = x0 - (d + 1);
#pragma unroll
    for (int c = 0; c < GPT * CPG; ++c) {
      r[c][3] = r[c][2];
      r[c][2] = r[c][1];
      r[c][1] = r[c][0];
      r[c][0] = (xn >= 0 && xn < W) ? __ldg(Rb + c * plane + xn) : 0.f;
    }
  }
}

// generic (any channels-per-group) fallback: one thread per output element group of 4 x
__global__ void __launch_bounds__(256) gwc_volume_generic_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                                                 float* __restrict__ V, int C, int H, int W, int D, int G,
                                                                 long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long t = i;
  const int x = t % W;
  t /= W;
  const int y = t % H;
  t /= H;
  const int d = t % D;
  t /= D;
  const int g = t % G;
  const int b = (int)(t / G);
  const int cpg = C / G;
  float s = 0.f;
  if (x >= d) {
    const long long plane = (long long)H * W;
    const float* Lp = L + ((long long)b * C + (long long)g * cpg) * plane + (long long)y * W + x;
    const float* Rp = R + ((long long)b * C + (long long)g * cpg) * plane + (long long)y * W + x - d;
    for (int k = 0; k < cpg; ++k) s = __fadd_rn(s, __fmul_rn(__ldg(Lp + k * plane), __ldg(Rp + k * plane)));
    s = s / (float)cpg;
  }
  V[i] = s;
}

// x / (||x||_2 + 1e-5) over channels, per pixel (norm_correlation, submodule.py:187-189)
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ X, float* __restrict__ Y, int C,
                                                           long long plane, long long total_pixels) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pixels) return;
  const long long b = i / plane;
  const long long p = i - b * plane;
  const float* x = X + b * C * plane + p;
  float* y = Y + b * C * plane + p;
  float ss = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = __ldg(x + c * plane);
    ss = fmaf(v, v, ss);
  }
  const float den = sqrtf(ss) + 1e-5f;
  for (int c = 0; c < C; ++c) y[c * plane] = __ldg(x + c * plane) / den;
}

// V[b,0,d,y,x] = mean_c Ln[c,y,x]*Rn[c,y,x-d]   (x >= d, else 0)
template <int DT>  // disparities per thread
__global__ void __launch_bounds__(128) norm_corr_kernel(const float* __restrict__ Ln, const float* __restrict__ Rn,
                                                        float* __restrict__ V, int C, int H, int W, int D) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int d0 = (blockIdx.z % ceil_div_dev(D, DT)) * DT;
  const int b = blockIdx.z / ceil_div_dev(D, DT);
  if (x >= W) return;
  const long long plane = (long long)H * W;
  const float* lp = Ln + (long long)b * C * plane + (long long)y * W + x;
  const float* rp = Rn + (long long)b * C * plane + (long long)y * W;
  float acc[DT];
#pragma unroll
  for (int j = 0; j < DT; ++j) acc[j] = 0.f;
  for (int c = 0; c < C; ++c) {
    const float l = __ldg(lp + c * plane);
#pragma unroll
    for (int j = 0; j < DT; ++j) {
      const int xr = x - (d0 + j);
      if (xr >= 0) acc[j] = __fadd_rn(acc[j], __fmul_rn(l, __ldg(rp + c * plane + xr)));
    }
  }
#pragma unroll
  for (int j = 0; j < DT; ++j) {
    const int d = d0 + j;
    if (d < D) V[((long long)b * D + d) * plane + (long long)y * W + x] = (x >= d) ? acc[j] / (float)C : 0.f;
  }
}

}  // namespace esm

using namespace esm;

extern "C" int esm_gwc_volume_f32(const float* L, const float* R, float* V, int B, int C, int H, int W, int D, int G,
                                  void* stream) {
  ESM_REQUIRE(L && R && V, "gwc_volume: null pointer");
  ESM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0 && G > 0, "gwc_volume: empty shape");
  ESM_REQUIRE(C % G == 0, "gwc_volume: C (%d) not divisible by groups (%d)", C, G);  // submodule.py:145
  ESM_REQUIRE(B <= 65535, "gwc_volume: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int cpg = C / G;
  if (cpg == 2) {
    constexpr int GPT = 2;
    const int xg = ceil_div(W, 4);
    const long long positions = (long long)H * xg;
    const int gsets = ceil_div(G, GPT);
    const int warps = 8;
    dim3 grid((unsigned)ceil_div_ll(positions, 32), (unsigned)ceil_div(gsets, warps), (unsigned)B);
    gwc_volume_kernel<GPT, 2><<<grid, warps * 32, 0, st>>>(L, R, V, C, H, W, D, G, xg, positions);
  } else {
    const long long total = (long long)B * G * D * H * W;
    gwc_volume_generic_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(L, R, V, C, H, W, D, G, total);
  }
  return check_launch("gwc_volume");
}

extern "C" int esm_norm_corr_volume_f32(const float* L, const float* R, float* V, float* ws, int B, int C, int H, int W,
                                        int D, void* stream) {
  ESM_REQUIRE(L && R && V && ws, "norm_corr_volume: null pointer");
  ESM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "norm_corr_volume: empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  const long long plane = (long long)H * W;
  const long long px = plane * B;
  float* Ln = ws;
  float* Rn = ws + px * C;
  l2_normalize_kernel<<<(unsigned)ceil_div_ll(px, 256), 256, 0, st>>>(L, Ln, C, plane, px);
  l2_normalize_kernel<<<(unsigned)ceil_div_ll(px, 256), 256, 0, st>>>(R, Rn, C, plane, px);
  constexpr int DT = 4;
  ESM_REQUIRE((long long)B * ceil_div(D, DT) <= 65535 && H <= 65535, "norm_corr_volume: grid too large");
  dim3 grid((unsigned)ceil_div(W, 128), (unsigned)H, (unsigned)(B * ceil_div(D, DT)));
  norm_corr_kernel<DT><<<grid, 128, 0, st>>>(Ln, Rn, V, C, H, W, D);
  return check_launch("norm_corr_volume");
}

namespace esm {

// V[b, c, d, y, x] = L[b, c, y, x] (c < C, whole row -- the reference does not mask the left half, submodule.py:134)
//                  = R[b, c - C, y, x - d] for x >= d, else 0 (c >= C).  One thread per 4 consecutive x of one (c, d, y) row.
__global__ void __launch_bounds__(256) concat_volume_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ V,
                                                            int C, int H, int W, int D, int xg, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x0 = (int)(i % xg) * 4;
  long long t = i / xg;
  const int y = (int)(t % H);
  t /= H;
  const int d = (int)(t % D);
  t /= D;
  const int c = (int)(t % (2 * C));
  const long long b = t / (2 * C);
  const bool left = c < C;
  const float* src = (left ? L : R) + ((b * C + (left ? c : c - C)) * H + y) * (long long)W;
  float* dst = V + (((b * 2 * C + c) * D + d) * H + y) * (long long)W;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    if (x < W) dst[x] = left ? __ldg(src + x) : (x >= d ? __ldg(src + x - d) : 0.f);
  }
}

// V[b, g, d, y, x] = sum_{c in group g} (L[c, y, x] - R[c, y, x - d])^2 for x >= d, else 0
__global__ void __launch_bounds__(256) substract_volume_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ V,
                                                               int C, int H, int W, int D, int G, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  long long t = i / W;
  const int y = (int)(t % H);
  t /= H;
  const int d = (int)(t % D);
  t /= D;
  const int g = (int)(t % G);
  const long long b = t / G;
  const int cpg = C / G;
  float s = 0.f;
  if (x >= d) {
    const long long plane = (long long)H * W;
    const float* l = L + (b * C + (long long)g * cpg) * plane + (long long)y * W + x;
    const float* r = R + (b * C + (long long)g * cpg) * plane + (long long)y * W + x - d;
    for (int c = 0; c < cpg; ++c) {
      const float df = __fsub_rn(__ldg(l + c * plane), __ldg(r + c * plane));
      s = __fadd_rn(s, __fmul_rn(df, df));
    }
  }
  V[i] = s;
}

// build_gwc_volume_norm + groupwise_correlation_norm (submodule.py:163-184):
//   V[b, g, d, y, x] = mean_{c in g} (L[c, y, x] / (|L_g(y, x)|_2 + 1e-5)) * (R[c, y, x - d] / (|R_g(y, x - d)|_2 + 1e-5))   for x >= d, else 0
// with the norms over the channels of the group at that pixel.  Same operation order as the reference (divide, multiply, sum, divide by cpg).
__global__ void __launch_bounds__(256) gwc_volume_norm_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ V,
                                                              int C, int H, int W, int D, int G, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  long long t = i / W;
  const int y = (int)(t % H);
  t /= H;
  const int d = (int)(t % D);
  t /= D;
  const int g = (int)(t % G);
  const long long b = t / G;
  const int cpg = C / G;
  float s = 0.f;
  if (x >= d) {
    const long long plane = (long long)H * W;
    const float* l = L + (b * C + (long long)g * cpg) * plane + (long long)y * W + x;
    const float* r = R + (b * C + (long long)g * cpg) * plane + (long long)y * W + x - d;
    float nl = 0.f, nr = 0.f;
    for (int c = 0; c < cpg; ++c) {
      const float a = __ldg(l + c * plane), q = __ldg(r + c * plane);
      nl = __fadd_rn(nl, __fmul_rn(a, a));
      nr = __fadd_rn(nr, __fmul_rn(q, q));
    }
    nl = __fadd_rn(__fsqrt_rn(nl), 1e-5f);
    nr = __fadd_rn(__fsqrt_rn(nr), 1e-5f);
    for (int c = 0; c < cpg; ++c)
      s = __fadd_rn(s, __fmul_rn(__fdiv_rn(__ldg(l + c * plane), nl), __fdiv_rn(__ldg(r + c * plane), nr)));
    s = __fdiv_rn(s, (float)cpg);
  }
  V[i] = s;
}

}  // namespace esm

extern "C" int esm_gwc_volume_norm_f32(const float* L, const float* R, float* V, int B, int C, int H, int W, int D, int G, void* stream) {
  ESM_REQUIRE(L && R && V, "gwc_volume_norm: null pointer");
  ESM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0 && G > 0, "gwc_volume_norm: empty shape");
  ESM_REQUIRE(C % G == 0, "gwc_volume_norm: C (%d) not divisible by groups (%d)", C, G);  // submodule.py:165
  const long long total = (long long)B * G * D * H * W;
  ESM_REQUIRE(esm::ceil_div_ll(total, 256) < (1ll << 31), "gwc_volume_norm: grid too large");
  esm::gwc_volume_norm_kernel<<<(unsigned)esm::ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(L, R, V, C, H, W, D, G, total);
  return esm::check_launch("gwc_volume_norm");
}

extern "C" int esm_concat_volume_f32(const float* L, const float* R, float* V, int B, int C, int H, int W, int D, void* stream) {
  ESM_REQUIRE(L && R && V, "concat_volume: null pointer");
  ESM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "concat_volume: empty shape");
  const int xg = esm::ceil_div(W, 4);
  const long long total = (long long)B * 2 * C * D * H * xg;
  ESM_REQUIRE(esm::ceil_div_ll(total, 256) < (1ll << 31), "concat_volume: grid too large");
  esm::concat_volume_kernel<<<(unsigned)esm::ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(L, R, V, C, H, W, D, xg, total);
  return esm::check_launch("concat_volume");
}

extern "C" int esm_substract_volume_f32(const float* L, const float* R, float* V, int B, int C, int H, int W, int D, int G, void* stream) {
  ESM_REQUIRE(L && R && V, "substract_volume: null pointer");
  ESM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0 && G > 0, "substract_volume: empty shape");
  ESM_REQUIRE(C % G == 0, "substract_volume: C (%d) not divisible by groups (%d)", C, G);  // submodule.py:107
  const long long total = (long long)B * G * D * H * W;
  ESM_REQUIRE(esm::ceil_div_ll(total, 256) < (1ll << 31), "substract_volume: grid too large");
  esm::substract_volume_kernel<<<(unsigned)esm::ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(L, R, V, C, H, W, D, G, total);
  return esm::check_launch("substract_volume");
}
