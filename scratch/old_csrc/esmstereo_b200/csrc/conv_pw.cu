// Pointwise (1x1 / 1x1x1) convolutions as a streaming kernel: the k1 layers of the path -- aggregation.agg_0.0 /
// agg_1.0 after the skip concatenation (ESMStereo.py:152-157), up_refinement.agg_* (:211-216), the last layer of
// every disparity MLP (k1 with padding 1, :248-253), FMBlock's 1x1 (shufflemixer.py:126) and UpShuffle's 1x1 +
// PixelShuffle(2) + SiLU (ESMStereo.py:265-268) -- move 30-60 MB for a fraction of a GFLOP: they are HBM-bound, and
// both persistent conv engines ran them at 1.5 TB/s (latency per tile, not bandwidth: the 16 -> 64 UpShuffle layer took
// 62 us for 38 MB).  Here a thread owns ONE output pixel and CO output channels in registers, walks the input
// channels with coalesced 4-byte loads (a warp reads 128 contiguous bytes per channel; 8 channels in flight), takes the
// weights as broadcast 16-byte shared-memory reads (one LDS.128 per 4 FMAs), and stores each channel coalesced --
// PixelShuffle(2) pairs as 8-byte stores.  Plain grid (one small CTA per 128 pixels x channel tile): occupancy, not a
// software pipeline, hides the latency.
//
// The same kernel, with a tap loop, runs the 2D layers with 1 or 3 INPUT channels -- the image stems (3 -> 32 k3 s2 at
// full resolution, ESMStereo.py:528-533) and the first layers on the disparity map (1 -> 32 k5 / k3 s2, :191,:245) --
// which the channel-chunked engines pad to 8 input channels: 27 / 25 / 9 loads and 32 accumulators per pixel.
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <string.h>

namespace esm {

// activations of 4 values: the branch-free GELU / SiLU of the tensor-core epilogues (ulp-level error, 16 / 5 instructions
// instead of erff / expf + IEEE division), the generic out-of-line path for the rest
__device__ __forceinline__ void pw_act4(float (&rv)[4], int act) {
  if (act == ESM_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) rv[j] = tc_gelu(rv[j]);
  } else if (act == ESM_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) rv[j] = tc_silu(rv[j]);
  } else if (act != ESM_ACT_NONE) {
    const float4 r = apply_act4(make_float4(rv[0], rv[1], rv[2], rv[3]), act);
    rv[0] = r.x; rv[1] = r.y; rv[2] = r.z; rv[3] = r.w;
  }
}

struct PwK {
  esm_src_t src[3];
  int nsrc;
  int B, Cin, CinPad, CoutPad, Cout;
  int Din, Hin, Win;      // input extent
  int OD, OH, OW;         // output extent (= input + 2 * pad in h, w)
  int ph, pw;
  int KH, KW, S;          // taps and stride (1 x 1, stride 1 for the pointwise layers)
  const float* weight;    // fp32 pack [tap][CinPad][CoutPad]
  const float* scale;
  const float* shift;
  int act, act2, ps;
  const float* out_mul;
  long long omB, omC, omH;
  const float* residual;
  float out_scale;
  float* out;
  long long oB, oC, oD, oH;
  long long pixels;       // OD * OH * OW
  int cotiles;
};

constexpr int PW_THREADS = 128;

// CO output channels per thread (one channel tile per blockIdx.y)
template <int CO>
__global__ void __launch_bounds__(PW_THREADS) pw_conv_kernel(const __grid_constant__ PwK p) {
  extern __shared__ __align__(16) float s_w[];  // [tap][Cin][CO]: this tile's weights, then [2][CO] scale / shift
  const int co0 = blockIdx.y * CO;
  const int b = blockIdx.z;
  const int taps = p.KH * p.KW;
  const int nw = taps * p.Cin * CO;
  for (int i = threadIdx.x; i < nw; i += PW_THREADS) {
    const int r = i / CO, c = i - r * CO;  // CO is a compile-time constant
    int tap = 0, ci = r;
    if (taps > 1) {
      tap = r / p.Cin;
      ci = r - tap * p.Cin;
    }
    s_w[i] = (co0 + c < p.CoutPad) ? __ldg(p.weight + ((long long)tap * p.CinPad + ci) * p.CoutPad + co0 + c) : 0.f;
  }
  float* s_aff = s_w + nw;
  if (threadIdx.x < 2 * CO) {
    const int c = threadIdx.x % CO, co = co0 + c;
    const float* src = threadIdx.x < CO ? p.scale : p.shift;
    s_aff[threadIdx.x] = (src && co < p.Cout) ? __ldg(src + co) : (threadIdx.x < CO ? 1.f : 0.f);
  }
  __syncthreads();
  // 32-bit pixel decode (the host checks OD * OH * OW < 2^31)
  const unsigned pix = blockIdx.x * PW_THREADS + threadIdx.x;
  if (pix >= (unsigned)p.pixels) return;
  const int ox = (int)(pix % (unsigned)p.OW);
  const unsigned t = pix / (unsigned)p.OW;
  const int oy = (int)(t % (unsigned)p.OH), oz = (int)(t / (unsigned)p.OH);
  // accumulators as packed pairs: one FFMA2 (fma.rn.f32x2) per two channels -- the kernel is issue-bound on its FMAs
  float2 acc2[CO / 2];
#pragma unroll
  for (int c = 0; c < CO / 2; ++c) acc2[c] = make_float2(0.f, 0.f);
  if (taps > 1) {
    // few input channels, many taps: one source, every (tap, channel) is one load and CO FMAs
    const esm_src_t& sr = p.src[0];
    const float* ib = sr.ptr + (long long)b * sr.sB;
    const float* wrow = s_w;
    for (int kh = 0; kh < p.KH; ++kh) {
      const int iy = oy * p.S - p.ph + kh;
      for (int kw = 0; kw < p.KW; ++kw) {
        const int ix = ox * p.S - p.pw + kw;
        const bool in = (unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win;
        for (int ci = 0; ci < p.Cin; ++ci, wrow += CO) {
          const float x = in ? __ldg(ib + (long long)ci * sr.sC + (long long)iy * sr.sH + ix) : 0.f;
          const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
          for (int j = 0; j < CO / 4; ++j) {
            const float4 w = w4[j];
            ffma2(acc2[2 * j], make_float2(x, x), make_float2(w.x, w.y));
            ffma2(acc2[2 * j + 1], make_float2(x, x), make_float2(w.z, w.w));
          }
        }
      }
    }
  }
  const int iy = oy - p.ph, ix = ox - p.pw;
  const bool inside = taps == 1 && (unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win;  // k1 with padding: the border sees zeros
  if (inside) {
    int ci0 = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      const esm_src_t& sr = p.src[s];
      const float* ip = sr.ptr + (long long)b * sr.sB + (long long)oz * sr.sD + (long long)iy * sr.sH + ix;
      const long long sC = sr.sC;
      int c = 0;
      for (; c + 8 <= sr.C; c += 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = __ldg(ip + (c + u) * sC);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4* w4 = reinterpret_cast<const float4*>(s_w + (ci0 + c + u) * CO);
#pragma unroll
          for (int j = 0; j < CO / 4; ++j) {
            const float4 w = w4[j];
            ffma2(acc2[2 * j], make_float2(x[u], x[u]), make_float2(w.x, w.y));
            ffma2(acc2[2 * j + 1], make_float2(x[u], x[u]), make_float2(w.z, w.w));
          }
        }
      }
      for (; c < sr.C; ++c) {
        const float x = __ldg(ip + c * sC);
        const float4* w4 = reinterpret_cast<const float4*>(s_w + (ci0 + c) * CO);
#pragma unroll
        for (int j = 0; j < CO / 4; ++j) {
          const float4 w = w4[j];
          ffma2(acc2[2 * j], make_float2(x, x), make_float2(w.x, w.y));
          ffma2(acc2[2 * j + 1], make_float2(x, x), make_float2(w.z, w.w));
        }
      }
      ci0 += sr.C;
    }
  }
  // epilogue: affine -> act -> (x out_mul) -> (+ residual) -> act2 -> scale -> store
  const bool post = p.out_mul || p.residual;
  const float oscale = p.out_scale;
  const long long obase = (long long)b * p.oB + (long long)oz * p.oD + (long long)oy * p.oH + ox;
  float* ops = p.out + ((long long)b * p.oB + (long long)(co0 >> 2) * p.oC + (long long)(2 * oy) * p.oH + 2 * ox);  // PixelShuffle(2) target
  float* op = p.out + obase + (long long)co0 * p.oC;
#pragma unroll
  for (int c4 = 0; c4 < CO; c4 += 4) {
    const int co = co0 + c4;
    if (co >= p.Cout) break;
    float rv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = acc2[(c4 + j) / 2];
      rv[j] = fmaf((j & 1) ? a.y : a.x, s_aff[c4 + j], s_aff[CO + c4 + j]);
    }
    pw_act4(rv, p.act);
    if (p.ps == 2) {
      // channel co -> (co / 4, row 2y + (co / 2) % 2, column 2x + co % 2): the 4 channels of a unit are one 2 x 2 block
      pw_act4(rv, p.act2);
      *reinterpret_cast<float2*>(ops) = make_float2(rv[0] * oscale, rv[1] * oscale);
      *reinterpret_cast<float2*>(ops + p.oH) = make_float2(rv[2] * oscale, rv[3] * oscale);
      ops += p.oC;
      continue;
    }
    if (post) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (co + j < p.Cout) {
          if (p.out_mul) rv[j] *= __ldg(p.out_mul + (long long)b * p.omB + (long long)(co + j) * p.omC + (long long)oy * p.omH + ox);
          if (p.residual) rv[j] += __ldg(p.residual + obase + (long long)(co + j) * p.oC);
        }
      }
    }
    pw_act4(rv, p.act2);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (co + j < p.Cout) op[(long long)j * p.oC] = rv[j] * oscale;
    op += 4 * p.oC;
  }
}

static long long pw_launches = 0;

bool pw_conv_plan(const esm_conv_t* d, PwPlan* plan) {
  if (d->src_mode != ESM_SRC_TENSORS || d->transposed || d->in_mul) return false;
  const bool k1 = d->kd == 1 && d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pd == 0;
  // 2D layers on 1 or 3 input channels (image stems, first layers on the disparity map): direct tap loop
  const bool thin = d->kd == 1 && d->pd == 0 && d->Dout == 1 && d->Din == 1 && d->Cin <= 4 && d->nsrc == 1 && d->kh == d->kw && d->kh >= 3 &&
                    (d->stride == 1 || d->stride == 2) && !d->pixel_shuffle;
  if (!k1 && !thin) return false;
  if ((k1 && d->Cin < 4) || d->Cout < 4) return false;
  if (d->pixel_shuffle && !(d->pixel_shuffle == 2 && d->Cout % 4 == 0 && !d->out_mul && !d->residual && (d->oH % 2) == 0 &&
                            (reinterpret_cast<uintptr_t>(d->out) & 7) == 0 && d->Dout == 1))
    return false;
  if (d->B > 65535) return false;
  const int cop8 = round_up(d->Cout, 8);
  // channels per thread: 32 accumulators at most (two CTAs' worth of registers stay resident); wider layers re-read the
  // input once per channel tile, from L2
  plan->CO = cop8 <= 8 ? 8 : cop8 <= 16 ? 16 : cop8 <= 24 ? 24 : 32;
  plan->cotiles = ceil_div(d->Cout, plan->CO);
  if (plan->cotiles > 8) return false;
  plan->smem = ((size_t)d->kh * d->kw * d->Cin + 2) * plan->CO * sizeof(float);
  if ((long long)d->Dout * d->Hout * d->Wout >= (1ll << 31)) return false;
  return plan->smem <= 96 * 1024;
}

int pw_conv_launch(const esm_conv_t* d, const PwPlan& plan, cudaStream_t st) {
  PwK k;
  memset(&k, 0, sizeof(k));
  for (int i = 0; i < d->nsrc; ++i) k.src[i] = d->src[i];
  k.nsrc = d->nsrc;
  k.B = d->B;
  k.Cin = d->Cin;
  const TcgPack tp = tcg_pack_geom(d->Cout, d->Cin, 1, d->kh, d->kw, 0);
  k.CinPad = (d->Cin == 1 && d->Cout > 4) ? 1 : round_up(d->Cin, 8);  // pad_cin (conv.cu): single-channel inputs are not padded
  k.CoutPad = (int)(tp.offset / ((long long)k.CinPad * d->kh * d->kw));  // the fp32 pack is [tap][CinPad][CoutPad]
  k.KH = d->kh;
  k.KW = d->kw;
  k.S = d->stride;
  k.Cout = d->Cout;
  k.Din = d->Din;
  k.Hin = d->Hin;
  k.Win = d->Win;
  k.OD = d->Dout;
  k.OH = d->Hout;
  k.OW = d->Wout;
  k.ph = d->ph;
  k.pw = d->pw;
  k.weight = d->weight;
  k.scale = d->scale;
  k.shift = d->shift;
  k.act = d->act;
  k.act2 = d->act2;
  k.ps = d->pixel_shuffle;
  k.out_mul = d->out_mul;
  k.omH = d->Wout;
  k.omC = (long long)d->Hout * d->Wout;
  k.omB = k.omC * d->Cout;
  k.residual = d->residual;
  k.out_scale = d->out_scale;
  k.out = d->out;
  k.oB = d->oB;
  k.oC = d->oC;
  k.oD = d->oD;
  k.oH = d->oH;
  k.pixels = (long long)d->Dout * d->Hout * d->Wout;
  k.cotiles = plan.cotiles;
  const dim3 grid((unsigned)ceil_div_ll(k.pixels, PW_THREADS), (unsigned)plan.cotiles, (unsigned)d->B);
  void (*fn)(const PwK) = plan.CO == 8 ? pw_conv_kernel<8> : plan.CO == 16 ? pw_conv_kernel<16> : plan.CO == 24 ? pw_conv_kernel<24> : pw_conv_kernel<32>;
  if (plan.smem > 48 * 1024 && cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess)
    return check_launch("conv(pw, cudaFuncSetAttribute)");
  fn<<<grid, PW_THREADS, plan.smem, st>>>(k);
  ++pw_launches;
  return check_launch("conv(pw)");
}

}  // namespace esm

extern "C" long long esm_pw_conv_launches(void) { return esm::pw_launches; }
