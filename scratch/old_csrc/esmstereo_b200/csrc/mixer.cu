// ShuffleMixer feature-mixing layers (shufflemixer.py:23-132) as fused per-pixel kernels.
// The reference runs each SMLayer as ~30 small ops (two `rearrange` copies + 5 elementwise/reduce
// kernels per hand-rolled LayerNorm, chunk/cat, two 1x1 convs, a channel-shuffle copy, a depthwise
// conv); with C in {8,16} a whole pixel fits in registers, so each half of an SMLayer is ONE kernel
// that reads C planes and writes C planes (bandwidth bound: 8*C bytes per pixel).
#include "common.cuh"

namespace esm {

template <int C>
struct MlpSmem {
  float ln_w[C];
  float fc0_w[C * (C / 2)];  // [hidden=C][C/2]
  float fc0_b[C];
  float fc2_w[(C / 2) * C];  // [C/2][hidden=C]
  float fc2_b[C / 2];
};

template <int C>
__device__ __forceinline__ void load_mlp(MlpSmem<C>& s, const esm_mixer_mlp_t& m, int tid, int nt) {
  for (int i = tid; i < C; i += nt) {
    s.ln_w[i] = m.ln_w[i];
    s.fc0_b[i] = m.fc0_b[i];
  }
  for (int i = tid; i < C * (C / 2); i += nt) {
    s.fc0_w[i] = m.fc0_w[i];
    s.fc2_w[i] = m.fc2_w[i];
  }
  for (int i = tid; i < C / 2; i += nt) s.fc2_b[i] = m.fc2_b[i];
}

// u = shuffle8(cat(MLP(LN(t)[:C/2]), LN(t)[C/2:])) + t   (t in registers, result written back into t)
template <int C>
__device__ __forceinline__ void ln_mlp_shuffle_residual(float (&t)[C], const MlpSmem<C>& s) {
  constexpr int HALF = C / 2;
  float mu = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) mu += t[c];
  mu = mu / (float)C;
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float dlt = t[c] - mu;
    var = fmaf(dlt, dlt, var);
  }
  var = var / (float)C;
  const float den = sqrtf(var + 1e-5f);  // shufflemixer.py:60-62: (x - mu) / sqrt(sigma + 1e-5) * weight
  float y[C];
#pragma unroll
  for (int c = 0; c < C; ++c) y[c] = (t[c] - mu) / den * s.ln_w[c];
  float hdn[C];
#pragma unroll
  for (int j = 0; j < C; ++j) {
    float a = s.fc0_b[j];
#pragma unroll
    for (int i = 0; i < HALF; ++i) a = fmaf(s.fc0_w[j * HALF + i], y[i], a);
    hdn[j] = silu(a);
  }
  float u[C];
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    float a = s.fc2_b[i];
#pragma unroll
    for (int j = 0; j < C; ++j) a = fmaf(s.fc2_w[i * C + j], hdn[j], a);
    u[i] = a;
  }
#pragma unroll
  for (int i = HALF; i < C; ++i) u[i] = y[i];
  // 'b (g d) h w -> b (d g) h w', g = 8: input channel g*(C/8)+d goes to output channel d*8+g
  constexpr int DD = C / 8;
  float o[C];
#pragma unroll
  for (int g = 0; g < 8; ++g)
#pragma unroll
    for (int d = 0; d < DD; ++d) o[d * 8 + g] = u[g * DD + d];
#pragma unroll
  for (int c = 0; c < C; ++c) t[c] = o[c] + t[c];
}

template <int C>
__global__ void __launch_bounds__(256) sm_pointwise_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           long long plane, long long total, esm_mixer_mlp_t m,
                                                           const float* __restrict__ extra) {
  __shared__ MlpSmem<C> s;
  load_mlp<C>(s, m, threadIdx.x, blockDim.x);
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / plane;
  const long long p = i - b * plane;
  const long long base = b * C * plane + p;
  float t[C];
#pragma unroll
  for (int c = 0; c < C; ++c) t[c] = __ldg(x + base + c * plane);
  ln_mlp_shuffle_residual<C>(t, s);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v = t[c];
    if (extra) v += __ldg(extra + base + c * plane);
    y[base + c * plane] = v;
  }
}

constexpr int SP_TX = 32, SP_TY = 8;

// K = depthwise kernel size when known at compile time (7 on the ESMStereo path), 0 = runtime loops
template <int C, int K>
__global__ void __launch_bounds__(SP_TX * SP_TY) sm_spatial_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                   int H, int W, const float* __restrict__ dw_w,
                                                                   const float* __restrict__ dw_b, int k,
                                                                   esm_mixer_mlp_t m, const float* __restrict__ extra) {
  extern __shared__ __align__(16) float dyn[];
  __shared__ MlpSmem<C> s;
  const int tid = threadIdx.y * SP_TX + threadIdx.x;
  const int nt = SP_TX * SP_TY;
  load_mlp<C>(s, m, tid, nt);
  const int r = k / 2;
  const int TWp = SP_TX + k - 1, THp = SP_TY + k - 1;
  float* tile = dyn;                  // [C][THp][TWp]
  float* wsm = dyn + C * THp * TWp;   // [C][k*k]
  float* bsm = wsm + C * k * k;       // [C]
  for (int i = tid; i < C * k * k; i += nt) wsm[i] = dw_w[i];
  for (int i = tid; i < C; i += nt) bsm[i] = dw_b[i];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * SP_TX - r, y0 = blockIdx.y * SP_TY - r;
  const long long plane = (long long)H * W;
  const float* xb = x + (long long)b * C * plane;
  for (int i = tid; i < C * THp * TWp; i += nt) {
    const int c = i / (THp * TWp);
    const int rem = i - c * THp * TWp;
    const int ty = rem / TWp, tx = rem - ty * TWp;
    const int gy = y0 + ty, gx = x0 + tx;
    tile[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(xb + c * plane + (long long)gy * W + gx) : 0.f;
  }
  __syncthreads();
  const int px = blockIdx.x * SP_TX + threadIdx.x, py = blockIdx.y * SP_TY + threadIdx.y;
  if (px >= W || py >= H) return;
  float t[C];
  if (K > 0) {
    // fully unrolled taps, two channels in flight: the serial fmaf chain of the runtime loop was
    // latency-bound (34 us per launch at 96x312x16)
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      float a0 = bsm[c], a1 = bsm[c + 1];
      const float* tp0 = tile + (c * THp + threadIdx.y) * TWp + threadIdx.x;
      const float* tp1 = tp0 + THp * TWp;
      const float* wp0 = wsm + c * K * K;
      const float* wp1 = wp0 + K * K;
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          a0 = fmaf(wp0[ky * K + kx], tp0[ky * TWp + kx], a0);
          a1 = fmaf(wp1[ky * K + kx], tp1[ky * TWp + kx], a1);
        }
      t[c] = a0;
      t[c + 1] = a1;
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float a = bsm[c];
      const float* tp = tile + (c * THp + threadIdx.y) * TWp + threadIdx.x;
      const float* wp = wsm + c * k * k;
      for (int ky = 0; ky < k; ++ky)
        for (int kx = 0; kx < k; ++kx) a = fmaf(wp[ky * k + kx], tp[ky * TWp + kx], a);
      t[c] = a;
    }
  }
  ln_mlp_shuffle_residual<C>(t, s);
  const long long base = (long long)b * C * plane + (long long)py * W + px;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v = t[c];
    if (extra) v += __ldg(extra + base + c * plane);
    y[base + c * plane] = v;
  }
}


// ---- whole SMLayer in one kernel (shufflemixer.py:97-112): u = shuffle(MLP1(LN1(x))) + x; t = depthwise7x7(u) + b;
// y = shuffle(MLP2(LN2(t))) + t [+ extra].  A CTA owns a 32 x 8 output tile: phase A runs the pointwise half on the
// tile + 3-pixel halo (the halo pixels are recomputed, 532 pixels for 256 outputs: that half is ~450 instructions per
// pixel) and leaves it in shared memory -- zeros outside the image, which is the depthwise conv's padding -- phase B
// gives every thread two horizontally adjacent pixels: one LDS.64 row segment and one padded weight row (2 x LDS.128)
// feed 14 FMAs, against two loads per FMA in sm_spatial_kernel (LDS-bound: 20 us per launch at 96 x 312 x 16).
constexpr int SL_TW = 32, SL_TH = 8, SL_K = 7, SL_R = SL_K / 2;
constexpr int SL_PW = SL_TW + SL_K - 1, SL_PH = SL_TH + SL_K - 1;  // 38 x 14

template <int C>
__global__ void __launch_bounds__(SL_TW / 2 * SL_TH) sm_layer_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W,
                                                                      esm_mixer_mlp_t m1, const float* __restrict__ dw_w,
                                                                      const float* __restrict__ dw_b, esm_mixer_mlp_t m2,
                                                                      const float* __restrict__ extra) {
  __shared__ MlpSmem<C> s1, s2;
  __shared__ __align__(16) float s_w[C][SL_K][8];  // depthwise rows padded to 8 floats
  __shared__ float s_b[C];
  __shared__ __align__(16) float tile[C][SL_PH][SL_PW];
  const int tid = threadIdx.x, nt = SL_TW / 2 * SL_TH;
  load_mlp<C>(s1, m1, tid, nt);
  load_mlp<C>(s2, m2, tid, nt);
  for (int i = tid; i < C * SL_K * 8; i += nt) {
    const int kx = i & 7, ky = (i >> 3) % SL_K, c = i / (8 * SL_K);
    s_w[c][ky][kx] = kx < SL_K ? dw_w[(c * SL_K + ky) * SL_K + kx] : 0.f;
  }
  for (int i = tid; i < C; i += nt) s_b[i] = dw_b[i];
  __syncthreads();
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * SL_TW - SL_R, y0 = blockIdx.y * SL_TH - SL_R;
  const long long plane = (long long)H * W;
  const float* xb = x + (long long)b * C * plane;
  // phase A: pointwise half on the haloed tile
  for (int i = tid; i < SL_PH * SL_PW; i += nt) {
    const int ty = i / SL_PW, tx = i - ty * SL_PW;
    const int gy = y0 + ty, gx = x0 + tx;
    float t[C];
    if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
      const float* p = xb + (long long)gy * W + gx;
#pragma unroll
      for (int c = 0; c < C; ++c) t[c] = __ldg(p + c * plane);
      ln_mlp_shuffle_residual<C>(t, s1);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) t[c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) tile[c][ty][tx] = t[c];
  }
  __syncthreads();
  // phase B: depthwise 7 x 7 on two adjacent pixels, then the second pointwise half
  const int lx = tid % (SL_TW / 2), ly = tid / (SL_TW / 2);
  const int px = blockIdx.x * SL_TW + 2 * lx, py = blockIdx.y * SL_TH + ly;
  float t0[C], t1[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float a0 = s_b[c], a1 = a0;
#pragma unroll
    for (int ky = 0; ky < SL_K; ++ky) {
      const float2* rp = reinterpret_cast<const float2*>(&tile[c][ly + ky][2 * lx]);
      const float2 r01 = rp[0], r23 = rp[1], r45 = rp[2], r67 = rp[3];
      const float r[8] = {r01.x, r01.y, r23.x, r23.y, r45.x, r45.y, r67.x, r67.y};
      const float4 wa = *reinterpret_cast<const float4*>(&s_w[c][ky][0]), wb = *reinterpret_cast<const float4*>(&s_w[c][ky][4]);
      const float w[7] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
      for (int kx = 0; kx < SL_K; ++kx) {
        a0 = fmaf(w[kx], r[kx], a0);
        a1 = fmaf(w[kx], r[kx + 1], a1);
      }
    }
    t0[c] = a0;
    t1[c] = a1;
  }
  if (py >= H || px >= W) return;
  const long long base = (long long)b * C * plane + (long long)py * W + px;
  ln_mlp_shuffle_residual<C>(t0, s2);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v = t0[c];
    if (extra) v += __ldg(extra + base + c * plane);
    y[base + c * plane] = v;
  }
  if (px + 1 < W) {
    ln_mlp_shuffle_residual<C>(t1, s2);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float v = t1[c];
      if (extra) v += __ldg(extra + base + 1 + c * plane);
      y[base + 1 + c * plane] = v;
    }
  }
}


static int check_mlp(const esm_mixer_mlp_t* m, int C) {
  ESM_REQUIRE(m && m->ln_w && m->fc0_w && m->fc0_b && m->fc2_w && m->fc2_b, "mixer: null MLP parameter");
  ESM_REQUIRE(m->hidden == C, "mixer: hidden (%d) must equal C (%d) (mlp_ratio 2 on C/2)", m->hidden, C);
  return ESM_OK;
}

}  // namespace esm

using namespace esm;

extern "C" int esm_sm_pointwise_f32(const float* x, float* y, int B, int C, int H, int W, const esm_mixer_mlp_t* mlp,
                                    const float* extra_residual, void* stream) {
  ESM_REQUIRE(x && y && B > 0 && H > 0 && W > 0, "sm_pointwise: null pointer or empty shape");
  ESM_REQUIRE(C == 8 || C == 16, "sm_pointwise: C must be 8 or 16 (got %d)", C);
  if (int e = check_mlp(mlp, C)) return e;
  const long long plane = (long long)H * W, total = plane * B;
  const unsigned grid = (unsigned)ceil_div_ll(total, 256);
  if (C == 16)
    sm_pointwise_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, plane, total, *mlp, extra_residual);
  else
    sm_pointwise_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, plane, total, *mlp, extra_residual);
  return check_launch("sm_pointwise");
}

extern "C" int esm_sm_spatial_f32(const float* x, float* y, int B, int C, int H, int W, const float* dw_w,
                                  const float* dw_b, int k, const esm_mixer_mlp_t* mlp, const float* extra_residual,
                                  void* stream) {
  ESM_REQUIRE(x && y && dw_w && dw_b && B > 0 && H > 0 && W > 0, "sm_spatial: null pointer or empty shape");
  ESM_REQUIRE(C == 8 || C == 16, "sm_spatial: C must be 8 or 16 (got %d)", C);
  ESM_REQUIRE(k >= 1 && k <= 9 && (k & 1), "sm_spatial: depthwise kernel must be odd and <= 9 (got %d)", k);
  ESM_REQUIRE(x != y, "sm_spatial: in-place not supported (halo reads)");
  if (int e = check_mlp(mlp, C)) return e;
  ESM_REQUIRE(B <= 65535 && ceil_div(H, SP_TY) <= 65535, "sm_spatial: grid too large");
  dim3 grid((unsigned)ceil_div(W, SP_TX), (unsigned)ceil_div(H, SP_TY), (unsigned)B), block(SP_TX, SP_TY);
  const size_t smem = ((size_t)C * (SP_TY + k - 1) * (SP_TX + k - 1) + (size_t)C * k * k + C) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 16 && k == 7)
    sm_spatial_kernel<16, 7><<<grid, block, smem, st>>>(x, y, H, W, dw_w, dw_b, k, *mlp, extra_residual);
  else if (C == 8 && k == 7)
    sm_spatial_kernel<8, 7><<<grid, block, smem, st>>>(x, y, H, W, dw_w, dw_b, k, *mlp, extra_residual);
  else if (C == 16)
    sm_spatial_kernel<16, 0><<<grid, block, smem, st>>>(x, y, H, W, dw_w, dw_b, k, *mlp, extra_residual);
  else
    sm_spatial_kernel<8, 0><<<grid, block, smem, st>>>(x, y, H, W, dw_w, dw_b, k, *mlp, extra_residual);
  return check_launch("sm_spatial");
}

extern "C" int esm_sm_layer_f32(const float* x, float* y, int B, int C, int H, int W, const esm_mixer_mlp_t* mlp1, const float* dw_w,
                                const float* dw_b, int k, const esm_mixer_mlp_t* mlp2, const float* extra_residual, void* stream) {
  ESM_REQUIRE(x && y && dw_w && dw_b && B > 0 && H > 0 && W > 0, "sm_layer: null pointer or empty shape");
  ESM_REQUIRE(C == 8 || C == 16, "sm_layer: C must be 8 or 16 (got %d)", C);
  ESM_REQUIRE(k == SL_K, "sm_layer: depthwise kernel must be 7 (got %d): use esm_sm_pointwise_f32 + esm_sm_spatial_f32", k);
  ESM_REQUIRE(x != y, "sm_layer: in-place not supported (halo reads)");
  if (int e = check_mlp(mlp1, C)) return e;
  if (int e = check_mlp(mlp2, C)) return e;
  ESM_REQUIRE(B <= 65535 && ceil_div(H, SL_TH) <= 65535, "sm_layer: grid too large");
  dim3 grid((unsigned)ceil_div(W, SL_TW), (unsigned)ceil_div(H, SL_TH), (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 16)
    sm_layer_kernel<16><<<grid, SL_TW / 2 * SL_TH, 0, st>>>(x, y, H, W, *mlp1, dw_w, dw_b, *mlp2, extra_residual);
  else
    sm_layer_kernel<8><<<grid, SL_TW / 2 * SL_TH, 0, st>>>(x, y, H, W, *mlp1, dw_w, dw_b, *mlp2, extra_residual);
  return check_launch("sm_layer");
}
