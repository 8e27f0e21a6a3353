// ShuffleMixer feature-mixing layers (shufflemixer.py:23-132) as fused per-pixel kernels.
// The reference runs each SMLayer as ~30 small ops (two `rearrange` copies + 5 elementwise/reduce
// kernels per hand-rolled LayerNorm, chunk/cat, two 1x1 convs, a channel-shuffle copy, a depthwise
// conv); with C in {8,16} a whole pixel fits in registers, so each half of an SMLayer is ONE kernel
// that reads C planes and writes C planes (bandwidth bound: 8*C bytes per pixel).
#include "common.cuh"
#include "tc_common.cuh"

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
  // shufflemixer.py:60-62: (x - mu) / sqrt(sigma + 1e-5) * weight.  One IEEE division per pixel and a multiply per
  // channel (within 1 ulp of C divisions, which were a sixth of this function's instructions).
  const float rden = 1.0f / sqrtf(var + 1e-5f);
  float y[C];
#pragma unroll
  for (int c = 0; c < C; ++c) y[c] = (t[c] - mu) * rden * s.ln_w[c];
  float hdn[C];
#pragma unroll
  for (int j = 0; j < C; ++j) {
    float a = s.fc0_b[j];
#pragma unroll
    for (int i = 0; i < HALF; ++i) a = fmaf(s.fc0_w[j * HALF + i], y[i], a);
    hdn[j] = tc_silu(a);  // ex2.approx + rcp.approx, relative error 3e-7 (expf + an IEEE division: ~25 instructions per value)
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
// y = shuffle(MLP2(LN2(t))) + t [+ extra].  A CTA of 256 threads owns a 28 x 8 output tile, in three phases:
//   A  the first pointwise half, one thread per pixel of the tile + 3-pixel halo (34 x 14 = 476 pixels: two nearly full
//      rounds of 256 threads), left in shared memory -- zeros outside the image, which is the depthwise conv's padding;
//   B  the depthwise 7 x 7 with one thread per (channel, 4-pixel strip): the channel's 49 weights stay in registers,
//      a window row is three aligned vector loads for 28 FMAs (sm_spatial_kernel spends two scalar loads per FMA and is
//      LDS-bound: 20 us per launch at 96 x 312 x 16), results to a second shared-memory tile;
//   C  the second pointwise half, one thread per output pixel.
// Every sum is taken in the order of sm_pointwise_kernel / sm_spatial_kernel: bit-identical to the two-launch form.
constexpr int SL_TW = 28, SL_TH = 8, SL_K = 7, SL_R = SL_K / 2;
constexpr int SL_PW = SL_TW + SL_K - 1, SL_PH = SL_TH + SL_K - 1;  // 34 x 14
constexpr int SL_PITCH = 36;                                        // row pitch of the haloed tile (16-byte aligned strips)
constexpr int SL_THREADS = 256;

template <int C>
struct SlSmem {
  MlpSmem<C> s1, s2;
  float4 w[C][SL_K][2];  // depthwise rows padded to 8 floats
  float b[C];
  float u[C][SL_PH][SL_PITCH];
  float t[C][SL_TH][SL_TW];
};

template <int C>
__global__ void __launch_bounds__(SL_THREADS) sm_layer_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W,
                                                              esm_mixer_mlp_t m1, const float* __restrict__ dw_w,
                                                              const float* __restrict__ dw_b, esm_mixer_mlp_t m2,
                                                              const float* __restrict__ extra) {
  extern __shared__ __align__(16) uint8_t sl_raw[];
  SlSmem<C>& sm = *reinterpret_cast<SlSmem<C>*>(sl_raw);
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  load_mlp<C>(sm.s1, m1, tid, SL_THREADS);
  load_mlp<C>(sm.s2, m2, tid, SL_THREADS);
  for (int i = tid; i < C * SL_K * 8; i += SL_THREADS) {
    const int kx = i & 7, ky = (i >> 3) % SL_K, c = i / (8 * SL_K);
    reinterpret_cast<float*>(&sm.w[c][ky][0])[kx] = kx < SL_K ? dw_w[(c * SL_K + ky) * SL_K + kx] : 0.f;
  }
  for (int i = tid; i < C; i += SL_THREADS) sm.b[i] = dw_b[i];
  __syncthreads();
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * SL_TW - SL_R, y0 = blockIdx.y * SL_TH - SL_R;
  const long long plane = (long long)H * W;
  const float* xb = x + (long long)b * C * plane;
  // phase A
  for (int i = tid; i < SL_PH * SL_PW; i += SL_THREADS) {
    const int ty = i / SL_PW, tx = i - ty * SL_PW;
    const int gy = y0 + ty, gx = x0 + tx;
    float t[C];
    if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
      const float* p = xb + (long long)gy * W + gx;
#pragma unroll
      for (int c = 0; c < C; ++c) t[c] = __ldg(p + c * plane);
      ln_mlp_shuffle_residual<C>(t, sm.s1);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) t[c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) sm.u[c][ty][tx] = t[c];
  }
  __syncthreads();
  // phase B
  {
    constexpr int TPC = SL_THREADS / C;               // threads per channel
    constexpr int NSTRIP = (SL_TW / 4) * SL_TH;       // 4-pixel strips per channel
    const int c = tid / TPC, sub = tid - c * TPC;
    float w[SL_K][8];
#pragma unroll
    for (int ky = 0; ky < SL_K; ++ky) {
      const float4 wa = sm.w[c][ky][0], wb = sm.w[c][ky][1];
      w[ky][0] = wa.x; w[ky][1] = wa.y; w[ky][2] = wa.z; w[ky][3] = wa.w;
      w[ky][4] = wb.x; w[ky][5] = wb.y; w[ky][6] = wb.z; w[ky][7] = 0.f;
    }
    const float bias = sm.b[c];
    for (int s = sub; s < NSTRIP; s += TPC) {
      const int sy = s / (SL_TW / 4), sx = s - sy * (SL_TW / 4);
      float a[4] = {bias, bias, bias, bias};
#pragma unroll
      for (int ky = 0; ky < SL_K; ++ky) {
        const float* rp = &sm.u[c][sy + ky][4 * sx];
        const float4 r0 = *reinterpret_cast<const float4*>(rp), r1 = *reinterpret_cast<const float4*>(rp + 4);
        const float2 r2 = *reinterpret_cast<const float2*>(rp + 8);
        const float r[10] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y};
#pragma unroll
        for (int kx = 0; kx < SL_K; ++kx)
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = fmaf(w[ky][kx], r[kx + j], a[j]);
      }
      *reinterpret_cast<float4*>(&sm.t[c][sy][4 * sx]) = make_float4(a[0], a[1], a[2], a[3]);
    }
  }
  __syncthreads();
  // phase C
  if (tid < SL_TW * SL_TH) {
    const int ly = tid / SL_TW, lx = tid - ly * SL_TW;
    const int px = blockIdx.x * SL_TW + lx, py = blockIdx.y * SL_TH + ly;
    if (py < H && px < W) {
      float t[C];
#pragma unroll
      for (int c = 0; c < C; ++c) t[c] = sm.t[c][ly][lx];
      ln_mlp_shuffle_residual<C>(t, sm.s2);
      const long long base = (long long)b * C * plane + (long long)py * W + px;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v = t[c];
        if (extra) v += __ldg(extra + base + c * plane);
        y[base + c * plane] = v;
      }
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
  if (C == 16) {
    if (cudaFuncSetAttribute((const void*)sm_layer_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SlSmem<16>)) != cudaSuccess)
      return check_launch("sm_layer(cudaFuncSetAttribute)");
    launch_k(pdl_enabled(32), sm_layer_kernel<16>, grid, dim3(SL_THREADS), sizeof(SlSmem<16>), st, x, y, H, W, *mlp1, dw_w, dw_b, *mlp2, extra_residual);
  } else {
    launch_k(pdl_enabled(32), sm_layer_kernel<8>, grid, dim3(SL_THREADS), sizeof(SlSmem<8>), st, x, y, H, W, *mlp1, dw_w, dw_b, *mlp2, extra_residual);
  }
  return check_launch("sm_layer");
}
