// The image-side 3 -> C k3 stride-2 convolutions at full resolution (feature.conv_stem of the backbone and stem_2[0],
// ESMStereo.py:49,529-533: BasicConv / Conv2d + BN + activation on the RGB pair) as one dedicated FP32 kernel.  With 3
// input channels there is nothing to stage by channel chunk: a CTA stages the 3-channel input tile once (even / odd
// columns apart, so that the stride-2 reads of a warp are conflict-free), a thread owns two vertically adjacent output
// pixels x all output channels as packed float2 accumulators (FFMA2), and each weight LDS.128 feeds 8 FMAs.
// 42 MB of traffic and 0.41 GFLOP per launch at KITTI shape: the generic engines took 36-43 us on it (one load per
// tap per thread, latency-bound); this one is bound by FP32 issue.
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace esm {

struct Stem3K {
  const float* x;
  long long sB, sC, sH;
  int B, H, W, Ho, Wo, Cout;
  const float* w;      // FP32 pack of esm_pack_conv_weight_f32: [tap = kh*3+kw][CinPad = 8][CoutPad]
  int CoutPad;
  const float* scale;
  const float* shift;
  int act, act2;
  float out_scale;
  float* out;
  long long oB, oC, oH;
};

constexpr int S3_GX = 16, S3_GY = 8;               // 2 x 2 pixel groups per CTA: 32 x 16 output pixels, x 2 channel halves = 256 threads
constexpr int S3_OW = 2 * S3_GX, S3_OH = 2 * S3_GY;
constexpr int S3_IW = 2 * S3_OW + 1, S3_IH = 2 * S3_OH + 1;   // input tile 65 x 33
constexpr int S3_HALF = S3_OW + 1;                 // even columns: 33, odd columns: 32 (+1 pad)

// CO = padded output channels (16 or 32); a thread owns a 2 x 2 block of output pixels x CO / 2 channels (4 x CO / 4
// float2 accumulators): one weight LDS.128 feeds 16 FMAs, one input LDS 2-8.
template <int CO>
__global__ void __launch_bounds__(2 * S3_GX * S3_GY) stem3_kernel(const Stem3K p) {
  constexpr int CH = CO / 2;  // channels per thread
  __shared__ __align__(16) float s_w[27 * CO];
  __shared__ float s_in[3][S3_IH][2][S3_HALF];
  __shared__ float s_aff[2][CO];
  const int tid = threadIdx.x, nt = 2 * S3_GX * S3_GY;
  for (int i = tid; i < 27 * CO; i += nt) {
    const int co = i % CO, t = i / CO;  // t = tap * 3 + ci
    const int tap = t / 3, ci = t % 3;
    s_w[i] = co < p.CoutPad ? __ldg(p.w + ((long long)tap * 8 + ci) * p.CoutPad + co) : 0.f;
  }
  for (int i = tid; i < 2 * CO; i += nt) {
    const int c = i % CO;
    const float* src = i < CO ? p.scale : p.shift;
    s_aff[i / CO][c] = (src && c < p.Cout) ? __ldg(src + c) : (i < CO ? 1.f : 0.f);
  }
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * S3_OW, oy0 = blockIdx.y * S3_OH;
  const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy0 - 1;
  const float* xb = p.x + (long long)b * p.sB;
  // one tile row per warp and step (8 warps x 13 steps cover the 3 x 33 rows), 3 column chunks per row: no divisions
  for (int row = tid >> 5; row < 3 * S3_IH; row += nt >> 5) {
    const int c = row / S3_IH, ty = row - c * S3_IH;
    const int gy = iy0 + ty;
    const bool row_ok = (unsigned)gy < (unsigned)p.H;
    const float* rp = xb + c * p.sC + (long long)gy * p.sH + ix0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int tx = (tid & 31) + 32 * k;
      if (tx < S3_IW) {
        const int gx = ix0 + tx;
        s_in[c][ty][tx & 1][tx >> 1] = (row_ok && (unsigned)gx < (unsigned)p.W) ? __ldg(rp + tx) : 0.f;
      }
    }
  }
  __syncthreads();
  const int lx = tid % S3_GX, ly = (tid / S3_GX) % S3_GY, ch = tid / (S3_GX * S3_GY);  // a warp: 2 rows of groups, one channel half
  float2 acc[4][CH / 2];  // [pixel (qy, qx)][channel pair]
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < CH / 2; ++c) acc[q][c] = make_float2(0.f, 0.f);
#pragma unroll 1  // 288 FFMA2 per channel: unrolling all three is 50 KB of code (instruction-cache bound, measured)
  for (int ci = 0; ci < 3; ++ci) {
    // the 5 x 5 input window of the 2 x 2 output block: tile rows 4 ly .. 4 ly + 4, tile columns 4 lx .. 4 lx + 4
    float v[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 5; ++c) v[r][c] = s_in[ci][4 * ly + r][c & 1][2 * lx + (c >> 1)];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* wp = reinterpret_cast<const float4*>(s_w + ((kh * 3 + kw) * 3 + ci) * CO + ch * CH);
        const float2 a00 = make_float2(v[kh][kw], v[kh][kw]), a01 = make_float2(v[kh][kw + 2], v[kh][kw + 2]);
        const float2 a10 = make_float2(v[kh + 2][kw], v[kh + 2][kw]), a11 = make_float2(v[kh + 2][kw + 2], v[kh + 2][kw + 2]);
#pragma unroll
        for (int c4 = 0; c4 < CH / 4; ++c4) {
          const float4 w4 = wp[c4];
          const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
          ffma2(acc[0][2 * c4], a00, w01); ffma2(acc[0][2 * c4 + 1], a00, w23);
          ffma2(acc[1][2 * c4], a01, w01); ffma2(acc[1][2 * c4 + 1], a01, w23);
          ffma2(acc[2][2 * c4], a10, w01); ffma2(acc[2][2 * c4 + 1], a10, w23);
          ffma2(acc[3][2 * c4], a11, w01); ffma2(acc[3][2 * c4 + 1], a11, w23);
        }
      }
  }
  const int ox = ox0 + 2 * lx;
  if (ox >= p.Wo) return;
  const int act = p.act, act2 = p.act2;
  const bool pair_ok = ox + 1 < p.Wo && (p.oH % 2) == 0 && (p.oC % 2) == 0 && (p.oB % 2) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 7) == 0;
#pragma unroll
  for (int qy = 0; qy < 2; ++qy) {
    const int oy = oy0 + 2 * ly + qy;
    if (oy >= p.Ho) continue;
    float v[2][CH];
#pragma unroll
    for (int qx = 0; qx < 2; ++qx)
#pragma unroll
      for (int c = 0; c < CH / 2; ++c) {
        const int cc = ch * CH + 2 * c;
        v[qx][2 * c] = fmaf(acc[qy * 2 + qx][c].x, s_aff[0][cc], s_aff[1][cc]);
        v[qx][2 * c + 1] = fmaf(acc[qy * 2 + qx][c].y, s_aff[0][cc + 1], s_aff[1][cc + 1]);
      }
    float* vf = &v[0][0];
    // one branch per activation, not per value; GELU through the branch-free erfc fit of the tensor-core epilogues
    // (absolute error <= 6e-8, tc_common.cuh) -- erff on every value cost more than the convolution itself
    if (act == ESM_ACT_GELU) {
#pragma unroll
      for (int c = 0; c < 2 * CH; ++c) vf[c] = tc_gelu(vf[c]);
    } else if (act == ESM_ACT_RELU6) {
#pragma unroll
      for (int c = 0; c < 2 * CH; ++c) vf[c] = fminf(fmaxf(vf[c], 0.f), 6.f);
    } else if (act == ESM_ACT_RELU) {
#pragma unroll
      for (int c = 0; c < 2 * CH; ++c) vf[c] = fmaxf(vf[c], 0.f);
    } else if (act == ESM_ACT_SILU) {
#pragma unroll
      for (int c = 0; c < 2 * CH; ++c) vf[c] = tc_silu(vf[c]);
    } else if (act != ESM_ACT_NONE) {
#pragma unroll
      for (int c = 0; c < 2 * CH; c += 4) {
        const float4 t4 = apply_act4(make_float4(vf[c], vf[c + 1], vf[c + 2], vf[c + 3]), act);
        vf[c] = t4.x; vf[c + 1] = t4.y; vf[c + 2] = t4.z; vf[c + 3] = t4.w;
      }
    }
    if (act2 == ESM_ACT_RELU6) {
#pragma unroll
      for (int c = 0; c < 2 * CH; ++c) vf[c] = fminf(fmaxf(vf[c], 0.f), 6.f);
    } else if (act2 != ESM_ACT_NONE) {
#pragma unroll
      for (int c = 0; c < 2 * CH; c += 4) {
        const float4 t4 = apply_act4(make_float4(vf[c], vf[c + 1], vf[c + 2], vf[c + 3]), act2);
        vf[c] = t4.x; vf[c + 1] = t4.y; vf[c + 2] = t4.z; vf[c + 3] = t4.w;
      }
    }
    float* o = p.out + (long long)b * p.oB + (long long)(ch * CH) * p.oC + (long long)oy * p.oH + ox;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (ch * CH + c >= p.Cout) break;
      if (pair_ok) {
        *reinterpret_cast<float2*>(o + (long long)c * p.oC) = make_float2(v[0][c] * p.out_scale, v[1][c] * p.out_scale);
      } else {
        o[(long long)c * p.oC] = v[0][c] * p.out_scale;
        if (ox + 1 < p.Wo) o[(long long)c * p.oC + 1] = v[1][c] * p.out_scale;
      }
    }
  }
}

bool stem3_eligible(const esm_conv_t* d) {
  return d->src_mode == ESM_SRC_TENSORS && d->nsrc == 1 && d->Cin == 3 && !d->transposed && d->stride == 2 && d->kd == 1 && d->kh == 3 &&
         d->kw == 3 && d->pd == 0 && d->ph == 1 && d->pw == 1 && d->Din == 1 && d->Dout == 1 && d->Cout <= 32 && !d->in_mul && !d->out_mul &&
         !d->residual && !d->pixel_shuffle && d->B <= 65535 && d->Hout * (long long)d->Wout >= 1024;
}

int stem3_launch(const esm_conv_t* d, cudaStream_t st) {
  Stem3K k;
  k.x = d->src[0].ptr;
  k.sB = d->src[0].sB; k.sC = d->src[0].sC; k.sH = d->src[0].sH;
  k.B = d->B; k.H = d->Hin; k.W = d->Win; k.Ho = d->Hout; k.Wo = d->Wout; k.Cout = d->Cout;
  k.w = d->weight;
  k.CoutPad = (int)(tcg_pack_geom(d->Cout, d->Cin, 1, 3, 3, 0).offset / (9ll * 8));
  k.scale = d->scale; k.shift = d->shift; k.act = d->act; k.act2 = d->act2; k.out_scale = d->out_scale;
  k.out = d->out; k.oB = d->oB; k.oC = d->oC; k.oH = d->oH;
  dim3 grid((unsigned)ceil_div(d->Wout, S3_OW), (unsigned)ceil_div(d->Hout, S3_OH), (unsigned)d->B), block(2 * S3_GX * S3_GY);
  if (d->Cout <= 16)
    stem3_kernel<16><<<grid, block, 0, st>>>(k);
  else
    stem3_kernel<32><<<grid, block, 0, st>>>(k);
  return check_launch("conv(stem3)");
}

}  // namespace esm
