// Direct (implicit-GEMM on the FP32 pipe) convolution family for the ESMStereo hot path.
//
// Why CUDA cores and not tcgen05 here (see DESIGN.md "conv engine"): the parity gate is fp32-exact
// top-2 indices, which single-pass TF32 cannot meet (hard part 2), and every
// GEMM on this path has N = Cout in {8..72}; with voxels on M the A operand (im2col rows) must be
// re-read from shared memory once per tap, so an SS-mode UMMA is bound by the 128 B/clk/SM shared
// memory port at ~N*32 MAC/clk/SM, i.e. below the FP32 pipe once the 3xTF32 split triples the
// traffic.  The FP32 pipe with packed FFMA2 was measured at 67 TFLOP/s in this exact inner loop.
//
// One kernel template covers conv k1/k3/k5 stride 1, k3 stride 2, and ConvTranspose k4 s2 p1
// (as 4 / 8 sub-pixel phase convolutions with a 2-tap kernel per dimension), in 2D and 3D, with:
//   * up to 3 channel-concatenated strided sources (replaces torch.cat and the crop-to-skip slices),
//   * or the group-wise correlation volume generated on the fly from left/right features,
//   * folded BatchNorm/bias affine, activation, broadcast multiply, residual add, second activation,
//   * PixelShuffle store.
// Each thread owns 4 consecutive output voxels along W x COG output channels (packed as float2
// pairs for FFMA2); a CTA owns a TD x TH x 4*TWG brick of voxels and ALL output channels, stages
// CK input channels of the brick + halo in shared memory together with the matching weights.
#include "common.cuh"

#include <string.h>

#include <type_traits>

namespace esm {

struct ConvK {
  esm_src_t src[3];
  int nsrc, src_mode, cpg;
  const float* in_mul;
  long long imB, imC, imH;
  int B, Cin, Din, Hin, Win;
  int OD, OH, OW;       // real output extent
  int Cout, CinPad, CoutPad;
  int KD, KH;           // taps per CTA pass in d / h (KW is a template parameter)
  int pd, ph, pw;       // in = j*S - p + tap
  int transposed, phases_d;
  const float* weight;
  long long phase_stride;  // packed weight elements per phase
  const float* scale;
  const float* shift;
  int act, act2;
  const float* out_mul;
  long long omB, omC, omH;
  const float* residual;
  float out_scale;
  int ps;
  float* out;
  long long oB, oC, oD, oH;
  // tiling
  int TWG, TH, TD, slots, nthreads;
  int ID, IH, IWP;
  int tilesW, tilesH, tilesD;
  int cosplit, COP;  // output channels are split over `cosplit` CTAs of COP (padded) channels each
};

template <int KW, int S, int COG, int CK>
__global__ void __launch_bounds__(320, 2) conv_kernel(const __grid_constant__ ConvK p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NV = 4;
  constexpr int XN = (NV - 1) * S + KW;
  constexpr int XL = (XN + 3) / 4 * 4;
  constexpr int NP = 4;  // fill positions per thread per pass over a (channel, depth) plane

  const int tid = threadIdx.x;
  const int NT = p.nthreads;
  const int b = blockIdx.y;

  // ---- phase (transposed conv) ----
  int pz_d = 0, pz_h = 0, pz_w = 0;
  int pd = p.pd, ph = p.ph, pw = p.pw;
  int osd = 1, osh = 1, osw = 1;
  const float* wbase = p.weight;
  const int co_base = (blockIdx.z % p.cosplit) * p.COP;
  if (p.transposed) {
    const int z = blockIdx.z / p.cosplit;
    pz_w = z & 1;
    pz_h = (z >> 1) & 1;
    pz_d = (p.phases_d == 2) ? ((z >> 2) & 1) : 0;
    pw = 1 - pz_w;
    ph = 1 - pz_h;
    pd = (p.phases_d == 2) ? 1 - pz_d : 0;
    osw = 2;
    osh = 2;
    osd = (p.phases_d == 2) ? 2 : 1;
    wbase += (long long)z * p.phase_stride;
  }

  // ---- tile ----
  int t = blockIdx.x;
  const int tileW = t % p.tilesW;
  t /= p.tilesW;
  const int tileH = t % p.tilesH;
  const int tileD = t / p.tilesH;
  const int TW = p.TWG * NV;
  const int iw0 = tileW * TW * S - pw;
  const int ih0 = tileH * p.TH * S - ph;
  const int id0 = tileD * p.TD * S - pd;

  const int slot = tid % p.slots;
  const int cog = tid / p.slots;
  const int twg = slot % p.TWG;
  const int th = (slot / p.TWG) % p.TH;
  const int td = slot / (p.TWG * p.TH);

  const int ID = p.ID, IH = p.IH, IWP = p.IWP;
  const int plane = IH * IWP;
  const int chan_stride = ID * plane;
  const int COP = p.COP;
  const int taps = p.KD * p.KH * KW;
  float* s_in = smem;
  float* s_w = smem + CK * chan_stride;

  float2 acc[NV][COG / 2];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < COG / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);

  const float* xin = s_in + ((td * S) * IH + th * S) * IWP + twg * NV * S;
  const float* wthr = s_w + cog * COG;
  const bool compute_thread = tid < p.slots * (COP / COG);

  for (int c0 = 0; c0 < p.Cin; c0 += CK) {
    __syncthreads();
    // ---------------- stage CK input channels (brick + halo) ----------------
    for (int pb = 0; pb < plane; pb += NP * NT) {
      int hh[NP], ww[NP];
      bool ok[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int pos = pb + tid + k * NT;
        const int hy = pos / IWP;
        const int col = pos - hy * IWP;
        hh[k] = ih0 + hy;
        ww[k] = iw0 + col;
        ok[k] = (pos < plane) && (hh[k] >= 0) && (hh[k] < p.Hin) && (ww[k] >= 0) && (ww[k] < p.Win);
      }
      for (int c = 0; c < CK; ++c) {
        const int cc = c0 + c;
        const bool cvalid = cc < p.Cin;
        // locate the source tensor of channel cc (ESM_SRC_TENSORS)
        const float* sbase = nullptr;
        long long sD = 0, sH = 0;
        const float* rbase = nullptr;  // gwc: right features
        long long sC = 0;
        if (cvalid) {
          if (p.src_mode == ESM_SRC_GWC) {
            sbase = p.src[0].ptr + (long long)b * p.src[0].sB + (long long)cc * p.cpg * p.src[0].sC;
            rbase = p.src[1].ptr + (long long)b * p.src[1].sB + (long long)cc * p.cpg * p.src[1].sC;
            sH = p.src[0].sH;
            sC = p.src[0].sC;
          } else {
            int rel = cc, k = 0;
            while (k < p.nsrc - 1 && rel >= p.src[k].C) {
              rel -= p.src[k].C;
              ++k;
            }
            sbase = p.src[k].ptr + (long long)b * p.src[k].sB + (long long)rel * p.src[k].sC;
            sD = p.src[k].sD;
            sH = p.src[k].sH;
          }
        }
        const float* mbase = (p.in_mul && cvalid) ? p.in_mul + (long long)b * p.imB + (long long)cc * p.imC : nullptr;
        long long off[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) off[k] = (long long)hh[k] * sH + ww[k];
        for (int dz = 0; dz < ID; ++dz) {
          const int d = id0 + dz;
          const bool dvalid = cvalid && d >= 0 && d < p.Din;
          float* dst = s_in + (c * ID + dz) * plane + pb;
#pragma unroll
          for (int k = 0; k < NP; ++k) {
            const int pos = tid + k * NT;
            if (pb + pos < plane) {
              float v = 0.f;
              if (dvalid && ok[k]) {
                if (p.src_mode == ESM_SRC_GWC) {
                  const int wr = ww[k] - d;  // right-image column, submodule.py:156
                  if (wr >= 0) {
                    const long long o = off[k];
                    float s = 0.f;
                    for (int q = 0; q < p.cpg; ++q)  // un-contracted: matches (fea1*fea2).mean(2), submodule.py:147
                      s = __fadd_rn(s, __fmul_rn(__ldg(sbase + q * sC + o), __ldg(rbase + q * sC + o - d)));
                    v = s / (float)p.cpg;
                  }
                } else {
                  v = __ldg(sbase + (long long)d * sD + off[k]);
                }
                if (mbase) v *= __ldg(mbase + (long long)hh[k] * p.imH + ww[k]);
              }
              dst[pos] = v;
            }
          }
        }
      }
    }
    {
      // ---------------- stage the matching weights ----------------
      // rows of COP floats per (tap, channel); 16-byte copies, 8-byte ones for the COG=2 (Cout<=2) layout
      using wvec_t = typename std::conditional<COG >= 4, float4, float2>::type;
      constexpr int WV = sizeof(wvec_t) / sizeof(float);
      const int copv = COP / WV;
      const int row_v = CK * copv;
      const wvec_t* wsrc = reinterpret_cast<const wvec_t*>(wbase);
      wvec_t* wdst = reinterpret_cast<wvec_t*>(s_w);
      for (int i = tid; i < taps * row_v; i += NT) {
        const int tap = i / row_v;
        const int rr = i - tap * row_v;
        const int c = rr / copv;
        const int r = rr - c * copv;
        wdst[i] = __ldg(wsrc + ((long long)(tap * p.CinPad + c0 + c) * p.CoutPad + co_base) / WV + r);
      }
    }
    __syncthreads();
    // ---------------- FFMA2 inner product ----------------
    if (compute_thread) {
      for (int kd = 0; kd < p.KD; ++kd) {
        for (int kh = 0; kh < p.KH; ++kh) {
          const float* xr = xin + (kd * IH + kh) * IWP;
          const float* wr = wthr + ((kd * p.KH + kh) * KW) * CK * COP;
#pragma unroll
          for (int c = 0; c < CK; ++c) {
            float x[XL];
#pragma unroll
            for (int q = 0; q < XL / 4; ++q) {
              const float4 t4 = *reinterpret_cast<const float4*>(xr + c * chan_stride + q * 4);
              x[q * 4 + 0] = t4.x;
              x[q * 4 + 1] = t4.y;
              x[q * 4 + 2] = t4.z;
              x[q * 4 + 3] = t4.w;
            }
#pragma unroll
            for (int kw = 0; kw < KW; ++kw) {
              float2 w2[COG / 2];
              const float* wp = wr + (kw * CK + c) * COP;
              if (COG >= 4) {
#pragma unroll
                for (int q = 0; q < COG / 4; ++q) {
                  const float4 t4 = *reinterpret_cast<const float4*>(wp + q * 4);
                  w2[q * 2 + 0] = make_float2(t4.x, t4.y);
                  w2[q * 2 + 1] = make_float2(t4.z, t4.w);
                }
              } else {
                w2[0] = *reinterpret_cast<const float2*>(wp);
              }
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                const float xv = x[v * S + kw];
                const float2 xx = make_float2(xv, xv);
#pragma unroll
                for (int j = 0; j < COG / 2; ++j) ffma2(acc[v][j], xx, w2[j]);
              }
            }
          }
        }
      }
    }
  }

  // ---------------- epilogue ----------------
  if (!compute_thread) return;
  const int jd = tileD * p.TD + td;
  const int jh = tileH * p.TH + th;
  const int jw0 = tileW * TW + twg * NV;
  const int od = jd * osd + pz_d;
  const int oh = jh * osh + pz_h;
  if (od >= p.OD || oh >= p.OH) return;
#pragma unroll
  for (int j = 0; j < COG; ++j) {
    const int co = co_base + cog * COG + j;
    if (co >= p.Cout) break;
    const float sc = p.scale ? __ldg(p.scale + co) : 1.f;
    const float sh = p.shift ? __ldg(p.shift + co) : 0.f;
    float r[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float a = (j & 1) ? acc[v][j / 2].y : acc[v][j / 2].x;
      r[v] = apply_act(fmaf(a, sc, sh), p.act);
    }
    if (p.ps == 0) {
      const long long obase = (long long)b * p.oB + (long long)co * p.oC + (long long)od * p.oD + (long long)oh * p.oH;
      const float* om = p.out_mul ? p.out_mul + (long long)b * p.omB + (long long)co * p.omC + (long long)oh * p.omH : nullptr;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int ow = (jw0 + v) * osw + pz_w;
        if (ow < p.OW) {
          float y = r[v];
          if (om) y *= __ldg(om + ow);
          if (p.residual) y += __ldg(p.residual + obase + ow);
          y = apply_act(y, p.act2) * p.out_scale;
          r[v] = y;
        }
      }
      float* o = p.out + obase;
      const int ow0 = jw0 * osw + pz_w;
      if (osw == 1 && ow0 + NV <= p.OW && ((reinterpret_cast<uintptr_t>(o + ow0) & 15) == 0)) {
        *reinterpret_cast<float4*>(o + ow0) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int ow = (jw0 + v) * osw + pz_w;
          if (ow < p.OW) o[ow] = r[v];
        }
      }
    } else {
      // PixelShuffle(r): channel co -> (c, a, bb); out[c, oh*r + a, ow*r + bb]   (2D only)
      const int rr = p.ps;
      const int c = co / (rr * rr);
      const int a = (co / rr) % rr;
      const int bb = co % rr;
      float* o = p.out + (long long)b * p.oB + (long long)c * p.oC + (long long)(oh * rr + a) * p.oH;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int ow = jw0 + v;
        if (ow < p.OW) o[ow * rr + bb] = apply_act(r[v], p.act2) * p.out_scale;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// weight packing / BN folding
// ------------------------------------------------------------------------------------------
static int pad_cout(int Cout) { return Cout <= 2 ? 2 : round_up(Cout, 8); }
// single-channel inputs (disparity / confidence maps) get their own CK=1 instantiation instead of
// 8x zero padding; the (Cin=1, Cout<=2) corner keeps the padded form so weight rows stay 16 B wide
static int pad_cin(int Cin, int Cout) { return (Cin == 1 && Cout > 2) ? 1 : round_up(Cin, 8); }

__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int kd,
                                   int kh, int kw, int transposed, int CinPad, int CoutPad, int KD, int KH, int KW,
                                   int phases, int phases_d, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long r = i;
  const int co = r % CoutPad;
  r /= CoutPad;
  const int ci = r % CinPad;
  r /= CinPad;
  const int tw = r % KW;
  r /= KW;
  const int thh = r % KH;
  r /= KH;
  const int tdd = r % KD;
  r /= KD;
  const int z = (int)r;  // phase
  float v = 0.f;
  if (co < Cout && ci < Cin) {
    if (!transposed) {
      v = w[(((long long)(co * Cin + ci) * kd + tdd) * kh + thh) * kw + tw];
    } else {
      const int pzw = z & 1, pzh = (z >> 1) & 1, pzd = (phases_d == 2) ? ((z >> 2) & 1) : 0;
      const int kkw = 3 - pzw - 2 * tw;
      const int kkh = 3 - pzh - 2 * thh;
      const int kkd = (phases_d == 2) ? 3 - pzd - 2 * tdd : 0;
      v = w[(((long long)(ci * Cout + co) * kd + kkd) * kh + kkh) * kw + kkw];  // [Cin,Cout,k,k,k]
    }
  }
  out[i] = v;
}

__global__ void fold_bn_kernel(const float* g, const float* bta, const float* mean, const float* var,
                               const float* bias, float eps, int C, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  float sc = 1.f, sh = 0.f;
  if (g) {
    sc = g[i] / sqrtf(var[i] + eps);
    sh = bta[i] - mean[i] * sc;
  }
  if (bias) sh += bias[i] * sc;
  scale[i] = sc;
  shift[i] = sh;
}

struct PackGeom {
  int KD, KH, KW, phases, phases_d, CinPad, CoutPad;
  long long per_phase;
};

static PackGeom pack_geom(int Cout, int Cin, int kd, int kh, int kw, int transposed) {
  PackGeom g;
  g.CinPad = pad_cin(Cin, Cout);
  g.CoutPad = pad_cout(Cout);
  if (transposed) {
    g.phases_d = (kd == 4) ? 2 : 1;
    g.KD = (kd == 4) ? 2 : 1;
    g.KH = 2;
    g.KW = 2;
    g.phases = 4 * g.phases_d;
  } else {
    g.phases_d = 1;
    g.KD = kd;
    g.KH = kh;
    g.KW = kw;
    g.phases = 1;
  }
  g.per_phase = (long long)g.KD * g.KH * g.KW * g.CinPad * g.CoutPad;
  return g;
}

// ------------------------------------------------------------------------------------------
// host-side tiling + dispatch
// ------------------------------------------------------------------------------------------
struct Tiling {
  int TWG, TH, TD, slots, IWP, ID, IH;
  size_t smem;
};

static bool choose_tiling(int Jw, int Jh, int Jd, int ncog, int KW, int KH, int KD, int S, int CK, int COP,
                          size_t smem_limit, Tiling* out) {
  int target = 256 / ncog;
  target = (target / 32) * 32;
  if (target < 32) target = 32;
  while (target * ncog > 320 && target > 32) target -= 32;
  if (target * ncog > 320) return false;
  double best = 1e30;
  bool found = false;
  for (int slots = target; slots >= 32; slots -= 32) {
    for (int TWG = 1; TWG <= 16; TWG *= 2) {
      if (slots % TWG) continue;
      const int R = slots / TWG;
      for (int TD = 1; TD <= R; ++TD) {
        if (R % TD) continue;
        const int TH = R / TD;
        if (Jd == 1 && TD != 1) continue;
        const int TW = TWG * 4;
        const double waste = (double)ceil_div(Jw, TW) * TW / Jw * ceil_div(Jh, TH) * TH / Jh * ceil_div(Jd, TD) * TD / Jd;
        const int ID = (TD - 1) * S + KD, IH = (TH - 1) * S + KH;
        const int XN = 3 * S + KW, XL = (XN + 3) / 4 * 4;
        int IWP = (TWG - 1) * 4 * S + XL;
        if (S == 1 && TWG < 8) {
          const int want = (4 * TWG) % 32;  // rows of an 8-lane LDS.128 phase land on distinct banks
          while (IWP % 32 != want) IWP += 4;
        }
        const double halo = (double)ID * IH * IWP / ((double)TD * TH * TW * S * S * (Jd == 1 ? 1 : S));
        const size_t smem = ((size_t)CK * ID * IH * IWP + (size_t)KD * KH * KW * CK * COP) * sizeof(float);
        if (smem > smem_limit) continue;
        // compute waste dominates; prefer bigger CTAs (fewer fills per FLOP) and small halos
        const double cost = waste * (1.0 + 0.05 * halo) * (1.0 + 8.0 / slots);
        if (cost < best) {
          best = cost;
          found = true;
          out->TWG = TWG;
          out->TH = TH;
          out->TD = TD;
          out->slots = slots;
          out->IWP = IWP;
          out->ID = ID;
          out->IH = IH;
          out->smem = smem;
        }
      }
    }
  }
  return found;
}

typedef void (*conv_fn_t)(const ConvK);

template <int KW, int S>
static conv_fn_t pick_cog_ck(int COG, int CK) {
  if (COG == 8 && CK == 8) return conv_kernel<KW, S, 8, 8>;
  if (COG == 8 && CK == 1) return conv_kernel<KW, S, 8, 1>;
  if (COG == 2 && CK == 8) return conv_kernel<KW, S, 2, 8>;
  return nullptr;
}

static conv_fn_t pick_kernel(int KW, int S, int COG, int CK) {
  if (S == 1) {
    if (KW == 1) return pick_cog_ck<1, 1>(COG, CK);
    if (KW == 2) return pick_cog_ck<2, 1>(COG, CK);
    if (KW == 3) return pick_cog_ck<3, 1>(COG, CK);
    if (KW == 5) return pick_cog_ck<5, 1>(COG, CK);
  } else if (S == 2 && KW == 3) {
    return pick_cog_ck<3, 2>(COG, CK);
  }
  return nullptr;
}

}  // namespace esm

using namespace esm;

extern "C" long long esm_packed_weight_elems(int Cout, int Cin, int kd, int kh, int kw, int transposed) {
  const PackGeom g = pack_geom(Cout, Cin, kd, kh, kw, transposed);
  return g.per_phase * g.phases;
}

extern "C" int esm_pack_conv_weight_f32(const float* w, float* packed, int Cout, int Cin, int kd, int kh, int kw,
                                        int transposed, void* stream) {
  ESM_REQUIRE(w && packed && Cout > 0 && Cin > 0, "pack_conv_weight: null pointer or empty shape");
  if (transposed)
    ESM_REQUIRE((kd == 4 || kd == 1) && kh == 4 && kw == 4, "pack_conv_weight: transposed conv must be k4 (got %d,%d,%d)",
                kd, kh, kw);
  const PackGeom g = pack_geom(Cout, Cin, kd, kh, kw, transposed);
  const long long total = g.per_phase * g.phases;
  const int threads = 256;
  pack_weight_kernel<<<(unsigned)ceil_div_ll(total, threads), threads, 0, (cudaStream_t)stream>>>(
      w, packed, Cout, Cin, kd, kh, kw, transposed, g.CinPad, g.CoutPad, g.KD, g.KH, g.KW, g.phases, g.phases_d, total);
  return check_launch("pack_conv_weight");
}

extern "C" int esm_fold_bn_f32(const float* gamma, const float* beta, const float* mean, const float* var,
                               const float* bias, float eps, int C, float* scale, float* shift, void* stream) {
  ESM_REQUIRE(scale && shift && C > 0, "fold_bn: null output");
  const bool any = gamma || beta || mean || var;
  ESM_REQUIRE(!any || (gamma && beta && mean && var), "fold_bn: BN tensors must be given together");
  fold_bn_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, bias, eps, C, scale, shift);
  return check_launch("fold_bn");
}

extern "C" int esm_conv_f32(const esm_conv_t* d, void* stream) {
  ESM_REQUIRE(d, "conv: null descriptor");
  ESM_REQUIRE(d->out && d->weight, "conv: null out/weight");
  ESM_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->Dout > 0 && d->Hout > 0 && d->Wout > 0, "conv: empty shape");
  ESM_REQUIRE(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  const int S = d->transposed ? 1 : d->stride;
  if (d->transposed) {
    ESM_REQUIRE((d->kd == 4 || d->kd == 1) && d->kh == 4 && d->kw == 4 && d->stride == 2 && d->ph == 1 && d->pw == 1 &&
                    d->pd == (d->kd == 4 ? 1 : 0),
                "conv: transposed conv supports k4 s2 p1 only");
    ESM_REQUIRE(d->Dout <= (d->kd == 4 ? 2 : 1) * d->Din && d->Hout <= 2 * d->Hin && d->Wout <= 2 * d->Win,
                "conv: transposed output larger than 2x input");
  } else {
    ESM_REQUIRE(d->Dout == (d->Din + 2 * d->pd - d->kd) / S + 1 && d->Hout == (d->Hin + 2 * d->ph - d->kh) / S + 1 &&
                    d->Wout == (d->Win + 2 * d->pw - d->kw) / S + 1,
                "conv: output extent does not match input/kernel/stride/padding");
  }
  ESM_REQUIRE(d->pixel_shuffle == 0 || (d->Dout == 1 && !d->transposed && d->Cout % (d->pixel_shuffle * d->pixel_shuffle) == 0),
              "conv: pixel_shuffle needs a 2D conv with Cout divisible by r*r");
  ESM_REQUIRE(d->pixel_shuffle == 0 || (!d->residual && !d->out_mul), "conv: pixel_shuffle excludes residual/out_mul");

  const PackGeom g = pack_geom(d->Cout, d->Cin, d->kd, d->kh, d->kw, d->transposed);
  ConvK k;
  memset(&k, 0, sizeof(k));
  k.src_mode = d->src_mode;
  if (d->src_mode == ESM_SRC_GWC) {
    ESM_REQUIRE(d->nsrc == 2 && d->gwc_groups == d->Cin && d->src[0].C == d->src[1].C && d->src[0].C % d->Cin == 0,
                "conv: ESM_SRC_GWC needs src[0]=left, src[1]=right with C %% groups == 0");
    k.cpg = d->src[0].C / d->Cin;
    ESM_REQUIRE(d->src[0].sC == d->src[1].sC && d->src[0].sH == d->src[1].sH, "conv: gwc sources must share strides");
  } else {
    ESM_REQUIRE(d->nsrc >= 1 && d->nsrc <= 3, "conv: nsrc must be 1..3");
    int csum = 0;
    for (int i = 0; i < d->nsrc; ++i) csum += d->src[i].C;
    ESM_REQUIRE(csum == d->Cin, "conv: source channels (%d) != Cin (%d)", csum, d->Cin);
  }
  for (int i = 0; i < d->nsrc; ++i) {
    ESM_REQUIRE(d->src[i].ptr, "conv: null source %d", i);
    k.src[i] = d->src[i];
  }
  k.nsrc = d->nsrc;
  k.in_mul = d->in_mul;
  k.imH = d->Win;
  k.imC = (long long)d->Hin * d->Win;
  k.imB = k.imC * d->Cin;
  k.B = d->B;
  k.Cin = d->Cin;
  k.Din = d->Din;
  k.Hin = d->Hin;
  k.Win = d->Win;
  k.OD = d->Dout;
  k.OH = d->Hout;
  k.OW = d->Wout;
  k.Cout = d->Cout;
  k.CinPad = g.CinPad;
  k.CoutPad = g.CoutPad;
  k.KD = g.KD;
  k.KH = g.KH;
  k.pd = d->pd;
  k.ph = d->ph;
  k.pw = d->pw;
  k.transposed = d->transposed;
  k.phases_d = g.phases_d;
  k.weight = d->weight;
  k.phase_stride = g.per_phase;
  k.scale = d->scale;
  k.shift = d->shift;
  k.act = d->act;
  k.act2 = d->act2;
  k.out_mul = d->out_mul;
  k.omH = d->Wout;
  k.omC = (long long)d->Hout * d->Wout;
  k.omB = k.omC * d->Cout;
  k.residual = d->residual;
  k.out_scale = d->out_scale;
  k.ps = d->pixel_shuffle;
  k.out = d->out;
  k.oB = d->oB;
  k.oC = d->oC;
  k.oD = d->oD;
  k.oH = d->oH;

  const int COG = g.CoutPad == 2 ? 2 : 8;
  const int CK = g.CinPad == 1 ? 1 : 8;
  // a CTA owns at most 80 output channels (10 channel groups x >=32 voxel slots <= 320 threads)
  int cosplit = 1;
  while (g.CoutPad / cosplit > 80 || g.CoutPad % cosplit || (g.CoutPad / cosplit) % COG) ++cosplit;
  const int COP = g.CoutPad / cosplit;
  const int ncog = COP / COG;
  // logical per-phase output extent
  const int Jw = d->transposed ? ceil_div(d->Wout, 2) : d->Wout;
  const int Jh = d->transposed ? ceil_div(d->Hout, 2) : d->Hout;
  const int Jd = (d->transposed && g.phases_d == 2) ? ceil_div(d->Dout, 2) : d->Dout;
  Tiling tl;
  // prefer tiles that leave room for 2 CTAs per SM (fill of one overlaps the math of the other)
  bool tiled = choose_tiling(Jw, Jh, Jd, ncog, g.KW, g.KH, g.KD, S, CK, COP, 110 * 1024, &tl) ||
               choose_tiling(Jw, Jh, Jd, ncog, g.KW, g.KH, g.KD, S, CK, COP, 220 * 1024, &tl);
  ESM_REQUIRE(tiled, "conv: no tiling for Cout=%d k=(%d,%d,%d)", d->Cout, d->kd, d->kh, d->kw);
  k.cosplit = cosplit;
  k.COP = COP;
  k.TWG = tl.TWG;
  k.TH = tl.TH;
  k.TD = tl.TD;
  k.slots = tl.slots;
  k.nthreads = tl.slots * ncog;
  k.ID = tl.ID;
  k.IH = tl.IH;
  k.IWP = tl.IWP;
  k.tilesW = ceil_div(Jw, tl.TWG * 4);
  k.tilesH = ceil_div(Jh, tl.TH);
  k.tilesD = ceil_div(Jd, tl.TD);

  conv_fn_t fn = pick_kernel(g.KW, S, COG, CK);
  ESM_REQUIRE(fn, "conv: unsupported kernel width %d / stride %d", g.KW, S);
  if (tl.smem > 48 * 1024) {
    if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem) != cudaSuccess)
      return check_launch("conv(cudaFuncSetAttribute)");
  }
  const long long ntiles = (long long)k.tilesW * k.tilesH * k.tilesD;
  ESM_REQUIRE(ntiles < (1ll << 31) && d->B <= 65535, "conv: grid too large");
  dim3 grid((unsigned)ntiles, (unsigned)d->B, (unsigned)(g.phases * cosplit));
  fn<<<grid, k.nthreads, tl.smem, (cudaStream_t)stream>>>(k);
  return check_launch("conv");
}
