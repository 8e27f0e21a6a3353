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

#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
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
  int phases, total_work, IWR;
};

// ---- cp.async helpers (LDGSTS): global -> shared without register staging; src_size 0 zero-fills ----
__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gsrc, bool pred) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
template <int BYTES>
__device__ __forceinline__ void cp_async_vec(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct TileCtx {
  int b, co_base, tileW, tileH, tileD;
  int pz_d, pz_h, pz_w;     // transposed-conv phase
  int pd, ph, pw;           // effective padding of this phase
  const float* wbase;
};

__device__ __forceinline__ TileCtx decode_work(const ConvK& p, int w) {
  TileCtx c;
  const int tiles = p.tilesW * p.tilesH * p.tilesD;
  int t = w % tiles;
  int r = w / tiles;
  c.tileW = t % p.tilesW;
  t /= p.tilesW;
  c.tileH = t % p.tilesH;
  c.tileD = t / p.tilesH;
  c.co_base = (r % p.cosplit) * p.COP;
  r /= p.cosplit;
  const int z = r % p.phases;
  c.b = r / p.phases;
  c.pz_d = c.pz_h = c.pz_w = 0;
  c.pd = p.pd;
  c.ph = p.ph;
  c.pw = p.pw;
  c.wbase = p.weight;
  if (p.transposed) {
    c.pz_w = z & 1;
    c.pz_h = (z >> 1) & 1;
    c.pz_d = (p.phases_d == 2) ? ((z >> 2) & 1) : 0;
    c.pw = 1 - c.pz_w;
    c.ph = 1 - c.pz_h;
    c.pd = (p.phases_d == 2) ? 1 - c.pz_d : 0;
    c.wbase += (long long)z * p.phase_stride;
  }
  return c;
}

// Persistent, double-buffered direct convolution.  Work items (tile, channel chunk) stream through a
// 2-stage shared-memory ring filled by cp.async: the loads of item i+1 are in flight while item i
// runs on the FP32 pipe.  GWC=true: the "input" voxels are group-wise correlations; the left/right
// feature rows of the next chunk are cp.async-staged and turned into the correlation tile
// smem -> smem (the D x H x W volume never exists in HBM).
template <int KW, int S, int COG, int CK, bool GWC>
__global__ void __launch_bounds__(320, 1) conv_kernel(const __grid_constant__ ConvK p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NV = 4;
  constexpr int XN = (NV - 1) * S + KW;
  constexpr int XL = (XN + 3) / 4 * 4;
  constexpr int NP = 4;  // fill positions per thread per pass over a plane

  const int tid = threadIdx.x;
  const int NT = p.nthreads;
  const int ID = p.ID, IH = p.IH, IWP = p.IWP;
  const int plane = IH * IWP;
  const int chan_stride = ID * plane;
  const int COP = p.COP;
  const int taps = p.KD * p.KH * KW;
  const int in_elems = CK * chan_stride;
  const int w_elems = taps * CK * COP;
  // smem carve-up: [in0][in1][w0][w1][w2 (GWC)][L staging][R staging]
  float* s_in0 = smem;
  float* s_w0 = smem + 2 * in_elems;
  const int IWR = p.IWR;                  // GWC: right staging row pitch
  float* s_L = s_w0 + 3 * w_elems;        // GWC only
  float* s_R = s_L + CK * p.cpg * plane;  // GWC only: [CK*cpg][IH][IWR]

  const int slot = tid % p.slots;
  const int cog = tid / p.slots;
  const int twg = slot % p.TWG;
  const int th = (slot / p.TWG) % p.TH;
  const int td = slot / (p.TWG * p.TH);
  const int TW = p.TWG * NV;
  const int xoff = ((td * S) * IH + th * S) * IWP + twg * NV * S;

  const int nch = (p.Cin + CK - 1) / CK;
  const int my_tiles = (p.total_work > (int)blockIdx.x) ? (p.total_work - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_items = my_tiles * nch;

  // ---------------- loaders ----------------
  auto load_weights = [&](int item, float* dst) {
    const int w = blockIdx.x + (item / nch) * gridDim.x;
    const int c0 = (item % nch) * CK;
    const TileCtx t = decode_work(p, w);
    constexpr int WB = (COG >= 4) ? 16 : 8;
    constexpr int WV = WB / 4;
    const int copv = COP / WV;
    const int row_v = CK * copv;
    for (int i = tid; i < taps * row_v; i += NT) {
      const int tap = i / row_v;
      const int rr = i - tap * row_v;
      const int c = rr / copv;
      const int r = rr - c * copv;
      cp_async_vec<WB>(dst + (long long)i * WV,
                       t.wbase + ((long long)(tap * p.CinPad + c0 + c) * p.CoutPad + t.co_base) + r * WV);
    }
  };

  auto load_inputs = [&](int item, float* dst) {  // ESM_SRC_TENSORS: brick + halo of CK channels
    const int w = blockIdx.x + (item / nch) * gridDim.x;
    const int c0 = (item % nch) * CK;
    const TileCtx t = decode_work(p, w);
    const int iw0 = t.tileW * TW * S - t.pw;
    const int ih0 = t.tileH * p.TH * S - t.ph;
    const int id0 = t.tileD * p.TD * S - t.pd;
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
#pragma unroll 1
      for (int c = 0; c < CK; ++c) {
        const int cc = c0 + c;
        const bool cvalid = cc < p.Cin;
        const float* sbase = p.src[0].ptr;
        long long sD = 0, sH = 0;
        if (cvalid) {
          int rel = cc;
          const esm_src_t* sp = &p.src[0];
          if (p.nsrc > 1 && rel >= p.src[0].C) {
            rel -= p.src[0].C;
            sp = &p.src[1];
            if (p.nsrc > 2 && rel >= p.src[1].C) {
              rel -= p.src[1].C;
              sp = &p.src[2];
            }
          }
          sbase = sp->ptr + (long long)t.b * sp->sB + (long long)rel * sp->sC;
          sD = sp->sD;
          sH = sp->sH;
        }
        long long off[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) off[k] = ok[k] ? (long long)hh[k] * sH + ww[k] : 0;
#pragma unroll 1
        for (int dz = 0; dz < ID; ++dz) {
          const int d = id0 + dz;
          const bool dvalid = cvalid && d >= 0 && d < p.Din;
          const float* sd = sbase + (dvalid ? (long long)d * sD : 0);
          float* drow = dst + (c * ID + dz) * plane + pb + tid;
#pragma unroll
          for (int k = 0; k < NP; ++k)
            if (pb + tid + k * NT < plane) cp_async_4(drow + k * NT, sd + off[k], dvalid && ok[k]);
        }
      }
    }
  };

  auto load_lr = [&](int item) {  // ESM_SRC_GWC: left rows [CK*cpg][IH][IWP], right rows [CK*cpg][IH][IWR]
    const int w = blockIdx.x + (item / nch) * gridDim.x;
    const int c0 = (item % nch) * CK;
    const TileCtx t = decode_work(p, w);
    const int iw0 = t.tileW * TW * S - t.pw;
    const int ih0 = t.tileH * p.TH * S - t.ph;
    const int id0 = t.tileD * p.TD * S - t.pd;
    const int nchan = CK * p.cpg;
    const int fc0 = c0 * p.cpg;  // first feature channel of this chunk of groups
    const int Cfeat = p.Cin * p.cpg;
    const float* Lb = p.src[0].ptr + (long long)t.b * p.src[0].sB;
    const float* Rb = p.src[1].ptr + (long long)t.b * p.src[1].sB;
    const long long sC = p.src[0].sC, sH = p.src[0].sH;
    for (int i = tid; i < nchan * plane; i += NT) {
      const int c = i / plane;
      const int rem = i - c * plane;
      const int hy = rem / IWP;
      const int col = rem - hy * IWP;
      const int h = ih0 + hy, x = iw0 + col;
      const bool ok = (fc0 + c < Cfeat) && h >= 0 && h < p.Hin && x >= 0 && x < p.Win;
      cp_async_4(s_L + i, Lb + (ok ? (long long)(fc0 + c) * sC + (long long)h * sH + x : 0), ok);
    }
    const int rw0 = iw0 - id0 - (ID - 1);  // right-image column of staging column 0
    const int rplane = IH * IWR;
    for (int i = tid; i < nchan * rplane; i += NT) {
      const int c = i / rplane;
      const int rem = i - c * rplane;
      const int hy = rem / IWR;
      const int col = rem - hy * IWR;
      const int h = ih0 + hy, x = rw0 + col;
      const bool ok = (fc0 + c < Cfeat) && h >= 0 && h < p.Hin && x >= 0 && x < p.Win;
      cp_async_4(s_R + i, Rb + (ok ? (long long)(fc0 + c) * sC + (long long)h * sH + x : 0), ok);
    }
  };

  auto build_volume = [&](int item, float* dst) {  // correlation tile from the staged rows (smem -> smem)
    const int w = blockIdx.x + (item / nch) * gridDim.x;
    const int c0 = (item % nch) * CK;
    const TileCtx t = decode_work(p, w);
    const int iw0 = t.tileW * TW * S - t.pw;
    const int ih0 = t.tileH * p.TH * S - t.ph;
    const int id0 = t.tileD * p.TD * S - t.pd;
    const float* mb = p.in_mul ? p.in_mul + (long long)t.b * p.imB : nullptr;
    const float inv = 1.0f / (float)p.cpg;
    const bool pow2 = (p.cpg & (p.cpg - 1)) == 0;
    for (int i = tid; i < CK * plane; i += NT) {
      const int g = i / plane;
      const int rem = i - g * plane;
      const int hy = rem / IWP;
      const int col = rem - hy * IWP;
      const int h = ih0 + hy, x = iw0 + col;
      const bool ok = (c0 + g < p.Cin) && h >= 0 && h < p.Hin && x >= 0 && x < p.Win;
      float m = 1.f;
      if (mb && ok) m = __ldg(mb + (long long)(c0 + g) * p.imC + (long long)h * p.imH + x);
      const float* lp = s_L + (g * p.cpg) * plane + rem;
      const float* rp = s_R + (g * p.cpg) * (IH * IWR) + hy * IWR + col + (ID - 1);
      for (int dz = 0; dz < ID; ++dz) {
        const int d = id0 + dz;
        float v = 0.f;
        if (ok && d >= 0 && d < p.Din && x - d >= 0) {
          float s = 0.f;
          for (int q = 0; q < p.cpg; ++q)  // un-contracted: (fea1*fea2).mean(2), submodule.py:147
            s = __fadd_rn(s, __fmul_rn(lp[q * plane], rp[q * IH * IWR - dz]));
          v = pow2 ? s * inv : s / (float)p.cpg;
          v *= m;
        }
        dst[(g * ID + dz) * plane + rem] = v;
      }
    }
  };

  auto scale_inputs = [&](int item, float* buf) {  // buf[c][dz][hy][col] *= in_mul[b, c0+c, 0, h, x]
    const int w = blockIdx.x + (item / nch) * gridDim.x;
    const int c0 = (item % nch) * CK;
    const TileCtx t = decode_work(p, w);
    const int iw0 = t.tileW * TW * S - t.pw;
    const int ih0 = t.tileH * p.TH * S - t.ph;
    const float* mb = p.in_mul + (long long)t.b * p.imB;
    for (int i = tid; i < CK * plane; i += NT) {
      const int c = i / plane;
      const int rem = i - c * plane;
      const int hy = rem / IWP;
      const int col = rem - hy * IWP;
      const int h = ih0 + hy, x = iw0 + col;
      if ((c0 + c < p.Cin) && h >= 0 && h < p.Hin && x >= 0 && x < p.Win) {
        const float m = __ldg(mb + (long long)(c0 + c) * p.imC + (long long)h * p.imH + x);
        for (int dz = 0; dz < ID; ++dz) buf[(c * ID + dz) * plane + rem] *= m;
      }
    }
  };

  float2 acc[NV][COG / 2];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < COG / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);

  // ---------------- prologue ----------------
  if (n_items > 0) {
    if (GWC) {
      load_weights(0, s_w0);
      load_lr(0);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      build_volume(0, s_in0);
      __syncthreads();
      if (n_items > 1) {
        load_weights(1, s_w0 + w_elems);
        load_lr(1);
      }
      cp_async_commit();
    } else {
      load_weights(0, s_w0);
      load_inputs(0, s_in0);
      cp_async_commit();
    }
  }

  for (int item = 0; item < n_items; ++item) {
    const float* s_in = s_in0 + (item & 1) * in_elems;
    const float* s_w;
    if (GWC) {
      s_w = s_w0 + (item % 3) * w_elems;
      cp_async_wait<0>();  // rows + weights of item+1 have landed
      __syncthreads();     // ...and every warp is done reading V[(item+1)&1] (FFMA2 of item-1)
      if (item + 1 < n_items) build_volume(item + 1, s_in0 + ((item + 1) & 1) * in_elems);
      __syncthreads();     // staging rows are free again
      if (item + 2 < n_items) {
        load_weights(item + 2, s_w0 + ((item + 2) % 3) * w_elems);
        load_lr(item + 2);
      }
      cp_async_commit();
    } else {
      s_w = s_w0 + (item & 1) * w_elems;
      if (item + 1 < n_items) {
        load_weights(item + 1, s_w0 + ((item + 1) & 1) * w_elems);
        load_inputs(item + 1, s_in0 + ((item + 1) & 1) * in_elems);
      }
      cp_async_commit();
      cp_async_wait<1>();  // everything but the group just committed -> item's data has landed
      __syncthreads();
      if (p.in_mul) {  // rare (unfused cv16 volume * att): scale the staged brick in place
        scale_inputs(item, s_in0 + (item & 1) * in_elems);
        __syncthreads();
      }
    }

    // ---------------- FFMA2 inner product ----------------
    {
      const float* xin = s_in + xoff;
      const float* wthr = s_w + cog * COG;
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

    // ---------------- epilogue (last channel chunk of a tile) ----------------
    if ((item % nch) == nch - 1) {
      const TileCtx t = decode_work(p, blockIdx.x + (item / nch) * gridDim.x);
      const int osd = (p.transposed && p.phases_d == 2) ? 2 : 1;
      const int osw = p.transposed ? 2 : 1;
      const int jd = t.tileD * p.TD + td;
      const int jh = t.tileH * p.TH + th;
      const int jw0 = t.tileW * TW + twg * NV;
      const int od = jd * osd + t.pz_d;
      const int oh = jh * osw + t.pz_h;
      const int b = t.b;
      if (od < p.OD && oh < p.OH) {
#pragma unroll
        for (int j = 0; j < COG; ++j) {
          const int co = t.co_base + cog * COG + j;
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
              const int ow = (jw0 + v) * osw + t.pz_w;
              if (ow < p.OW) {
                float y = r[v];
                if (om) y *= __ldg(om + ow);
                if (p.residual) y += __ldg(p.residual + obase + ow);
                y = apply_act(y, p.act2) * p.out_scale;
                r[v] = y;
              }
            }
            float* o = p.out + obase;
            const int ow0 = jw0 * osw + t.pz_w;
            if (osw == 1 && ow0 + NV <= p.OW && ((reinterpret_cast<uintptr_t>(o + ow0) & 15) == 0)) {
              *reinterpret_cast<float4*>(o + ow0) = make_float4(r[0], r[1], r[2], r[3]);
            } else {
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                const int ow = (jw0 + v) * osw + t.pz_w;
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
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < COG / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);
    }
    if (!GWC) __syncthreads();  // stage (item&1) may be refilled by the loads issued next iteration
  }
  cp_async_wait<0>();
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
  int TWG, TH, TD, slots, IWP, ID, IH, IWR;
  size_t smem;
};

static size_t conv_smem_bytes(int CK, int ID, int IH, int IWP, int taps, int COP, bool gwc, int cpg, int* IWR_out) {
  const size_t in_elems = (size_t)CK * ID * IH * IWP;
  const size_t w_elems = (size_t)taps * CK * COP;
  size_t total = 2 * in_elems + (gwc ? 3 : 2) * w_elems;
  int IWR = 0;
  if (gwc) {
    IWR = round_up(IWP + ID - 1, 4);
    total += (size_t)CK * cpg * IH * IWP + (size_t)CK * cpg * IH * IWR;
  }
  if (IWR_out) *IWR_out = IWR;
  return total * sizeof(float);
}

// Per-SM residency estimate: the kernels compile to <=128 registers, so at most 512 threads per SM;
// shared memory: 227 KB usable per SM, 1 KB reserved per CTA.
static int resident_ctas(int nthreads, size_t smem) {
  const int by_regs = 512 / nthreads;
  const int by_smem = (int)((227 * 1024) / (smem + 1024));
  int r = by_regs < by_smem ? by_regs : by_smem;
  return r > 8 ? 8 : r;
}

static bool choose_tiling(int Jw, int Jh, int Jd, int ncog, int KW, int KH, int KD, int S, int CK, int COP, bool gwc,
                          int cpg, Tiling* out, double* best_cost) {
  bool found = false;
  for (int slots = 32; slots * ncog <= 320 && slots <= 256; slots += 32) {
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
        int IWR = 0;
        const size_t smem = conv_smem_bytes(CK, ID, IH, IWP, KD * KH * KW, COP, gwc, cpg, &IWR);
        if (smem > 224 * 1024) continue;
        const int nthreads = slots * ncog;
        const int ctas = resident_ctas(nthreads, smem);
        if (ctas < 1) continue;
        const double halo = (double)ID * IH * IWP / ((double)TD * TH * TW * S * S * (Jd == 1 ? 1 : S));
        // FMA-pipe utilisation needs ~12+ resident warps per SM to cover LDS/FFMA2 latencies; below
        // that the cost grows quickly.  Then: wasted lanes, fill traffic (halo), barriers per FLOP
        // (small CTAs / small channel chunks sync more often).
        const double warps = (double)ctas * nthreads / 32.0;
        const double occ = warps >= 14.0 ? 1.0 : 14.0 / warps;
        const double cost = waste * occ * (1.0 + 0.05 * halo) * (1.0 + 8.0 / slots) * (CK >= 8 ? 1.0 : 1.04);
        if (cost < *best_cost) {
          *best_cost = cost;
          found = true;
          out->TWG = TWG;
          out->TH = TH;
          out->TD = TD;
          out->slots = slots;
          out->IWP = IWP;
          out->ID = ID;
          out->IH = IH;
          out->IWR = IWR;
          out->smem = smem;
        }
      }
    }
  }
  return found;
}

typedef void (*conv_fn_t)(const ConvK);

template <int KW, int S>
static conv_fn_t pick_cog_ck(int COG, int CK, bool gwc) {
  if (gwc) {
    if (!(COG == 8 && KW == 3 && S == 1)) return nullptr;
    return CK == 8 ? (conv_fn_t)conv_kernel<3, 1, 8, 8, true> : CK == 4 ? (conv_fn_t)conv_kernel<3, 1, 8, 4, true> : nullptr;
  }
  if (COG == 8 && CK == 8) return conv_kernel<KW, S, 8, 8, false>;
  if (COG == 8 && CK == 4) return conv_kernel<KW, S, 8, 4, false>;
  if (COG == 8 && CK == 1) return conv_kernel<KW, S, 8, 1, false>;
  if (COG == 2 && CK == 8) return conv_kernel<KW, S, 2, 8, false>;
  return nullptr;
}

static conv_fn_t pick_kernel(int KW, int S, int COG, int CK, bool gwc) {
  if (S == 1) {
    if (KW == 1) return pick_cog_ck<1, 1>(COG, CK, gwc);
    if (KW == 2) return pick_cog_ck<2, 1>(COG, CK, gwc);
    if (KW == 3) return pick_cog_ck<3, 1>(COG, CK, gwc);
    if (KW == 5) return pick_cog_ck<5, 1>(COG, CK, gwc);
  } else if (S == 2 && KW == 3) {
    return pick_cog_ck<3, 2>(COG, CK, gwc);
  }
  return nullptr;
}

// Launch plans are memoised per shape: the tiling search and the occupancy query cost ~50 us on the
// host, which matters in eager mode (under CUDA-graph replay the host never runs them).
struct PlanKey {
  int v[16];
  bool operator<(const PlanKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
struct Plan {
  Tiling tl;
  conv_fn_t fn;
  int cosplit, COP, COG, CK, blocks_per_sm;
};

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

  const bool gwc = d->src_mode == ESM_SRC_GWC;
  const int Jw = d->transposed ? ceil_div(d->Wout, 2) : d->Wout;
  const int Jh = d->transposed ? ceil_div(d->Hout, 2) : d->Hout;
  const int Jd = (d->transposed && g.phases_d == 2) ? ceil_div(d->Dout, 2) : d->Dout;

  static std::map<PlanKey, Plan> plans;
  static std::mutex plans_mu;
  static int num_sms = 0;
  PlanKey key = {{d->Cin, d->Cout, d->kd, d->kh, d->kw, d->stride, d->transposed, Jw, Jh, Jd, gwc ? k.cpg : 0, 0, 0, 0, 0, 0}};
  Plan plan;
  {
    std::lock_guard<std::mutex> lock(plans_mu);
    if (num_sms == 0) {
      int dev = 0;
      cudaDeviceProp prop;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        // no device: still run the validation / tiling below so shapes can be checked on a CPU box
        cudaGetLastError();
        num_sms = -1;
      } else {
        num_sms = prop.multiProcessorCount;
      }
    }
    auto it = plans.find(key);
    if (it == plans.end()) {
      Plan np;
      np.COG = g.CoutPad == 2 ? 2 : 8;
      np.CK = g.CinPad == 1 ? 1 : 8;
      // a CTA owns at most 80 output channels (10 channel groups x >=32 voxel slots <= 320 threads)
      int cosplit = 1;
      while (g.CoutPad / cosplit > 80 || g.CoutPad % cosplit || (g.CoutPad / cosplit) % np.COG) ++cosplit;
      np.cosplit = cosplit;
      np.COP = g.CoutPad / cosplit;
      const int ncog = np.COP / np.COG;
      // search tile shapes and channel-chunk depth (8 or 4: a shallower chunk halves the staged
      // brick, which buys resident warps on the wide 8-channel layers)
      double best = 1e30;
      bool tiled = choose_tiling(Jw, Jh, Jd, ncog, g.KW, g.KH, g.KD, S, np.CK, np.COP, gwc, k.cpg, &np.tl, &best);
      if (np.CK == 8 && np.COG == 8) {
        Tiling t4;
        if (choose_tiling(Jw, Jh, Jd, ncog, g.KW, g.KH, g.KD, S, 4, np.COP, gwc, k.cpg, &t4, &best)) {
          np.tl = t4;
          np.CK = 4;
          tiled = true;
        }
      }
      ESM_REQUIRE(tiled, "conv: no tiling for Cin=%d Cout=%d k=(%d,%d,%d)%s", d->Cin, d->Cout, d->kd, d->kh, d->kw,
                  gwc ? " +gwc" : "");
      np.fn = pick_kernel(g.KW, S, np.COG, np.CK, gwc);
      ESM_REQUIRE(np.fn, "conv: unsupported kernel width %d / stride %d%s", g.KW, S, gwc ? " with ESM_SRC_GWC" : "");
      np.blocks_per_sm = 1;
      if (num_sms > 0) {
        if (np.tl.smem > 48 * 1024 &&
            cudaFuncSetAttribute((const void*)np.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024) != cudaSuccess)
          return check_launch("conv(cudaFuncSetAttribute)");
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)np.fn, np.tl.slots * ncog, np.tl.smem) != cudaSuccess)
          return check_launch("conv(occupancy)");
        np.blocks_per_sm = occ > 0 ? occ : 1;
      }
      if (getenv("ESM_DEBUG_PLAN"))
        fprintf(stderr, "[esm plan] Cin=%d Cout=%d k=(%d,%d,%d) s=%d%s%s J=(%d,%d,%d): CK=%d COP=%d cosplit=%d tile=(%d,%d,%d) "
                "threads=%d smem=%zu KB ctas/SM(est)=%d occ(query)=%d\n",
                d->Cin, d->Cout, d->kd, d->kh, d->kw, d->stride, d->transposed ? " T" : "", gwc ? " gwc" : "", Jd, Jh, Jw, np.CK,
                np.COP, np.cosplit, np.tl.TD, np.tl.TH, np.tl.TWG * 4, np.tl.slots * ncog, np.tl.smem / 1024,
                resident_ctas(np.tl.slots * ncog, np.tl.smem), np.blocks_per_sm);
      it = plans.emplace(key, np).first;
    }
    plan = it->second;
  }
  const Tiling& tl = plan.tl;
  const int ncog = plan.COP / plan.COG;
  k.cosplit = plan.cosplit;
  k.COP = plan.COP;
  k.TWG = tl.TWG;
  k.TH = tl.TH;
  k.TD = tl.TD;
  k.slots = tl.slots;
  k.nthreads = tl.slots * ncog;
  k.ID = tl.ID;
  k.IH = tl.IH;
  k.IWP = tl.IWP;
  k.IWR = tl.IWR;
  k.tilesW = ceil_div(Jw, tl.TWG * 4);
  k.tilesH = ceil_div(Jh, tl.TH);
  k.tilesD = ceil_div(Jd, tl.TD);
  k.phases = g.phases;
  const long long total = (long long)k.tilesW * k.tilesH * k.tilesD * d->B * g.phases * plan.cosplit;
  ESM_REQUIRE(total < (1ll << 30), "conv: too many tiles");
  k.total_work = (int)total;
  if (num_sms <= 0) {
    set_error("conv: no CUDA device");
    return ESM_ERR_CUDA;
  }
  const long long resident = (long long)num_sms * plan.blocks_per_sm;
  const unsigned grid = (unsigned)(total < resident ? total : resident);  // persistent CTAs stride over the tiles
  plan.fn<<<grid, k.nthreads, tl.smem, (cudaStream_t)stream>>>(k);
  return check_launch("conv");
}
