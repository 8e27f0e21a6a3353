// Device side of the direct-convolution family (see conv.cu for the design notes).
#pragma once
#include "common.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

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
  int phases, total_work, IWR, IWL;  // GWC staging row pitches (right / left)
  int pzw_sel;  // transposed conv: -1 = all W phases in this launch, 0/1 = only that W phase (TMA launches)
  int nstages;  // ring depth of the TMA pipeline (2..4); cp.async and GWC paths use 2
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

// ---- TMA + mbarrier helpers (cp.async.bulk.tensor: hardware box copies with zero-filled halos) ----
struct __align__(64) ConvMaps {
  CUtensorMap src[3];  // per source: 5D (W,H,D,C,B) fp32, box (IWP, IH, ID, CK, 1); GWC: left / right rows
  CUtensorMap w;       // packed weights as 3D (CoutPad, CinPad, taps*phases), box (COP, CK, taps)
};

__device__ __forceinline__ unsigned smem_u32(const void* ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned done;
  do {
#ifdef CONV_WAIT_HINT_NS  // suspend-time hint as in tc_common.cuh: measured neutral on this engine (1.9132 vs 1.9133 ms), off
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"((unsigned)CONV_WAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#endif
  } while (!done);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int x0, int x1, int x2, int x3, int x4,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x0, int x1, int x2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x0), "r"(x1), "r"(x2)
      : "memory");
}

struct TileCtx {
  int b, z, co_base, tileW, tileH, tileD;
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
  int z = r % p.phases;
  c.b = r / p.phases;
  if (p.pzw_sel >= 0) z = (z << 1) | p.pzw_sel;  // this launch enumerates only (d,h) phases
  c.z = z;
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
// 2-stage shared-memory ring; the loads of the next items are in flight while item i runs on the
// FP32 pipe.  TMA=true: one elected thread issues cp.async.bulk.tensor box copies (brick + halo of CK
// channels, zero-filled out of bounds by the hardware, plus the weight slab) that complete on an
// mbarrier -- no per-element address math at all; needs 16-byte aligned pitches.  TMA=false: the
// same ring filled by 4-byte cp.async (any strides).  GWC=true: the "input" voxels are group-wise
// correlations; the left/right feature rows of the next chunk are staged (TMA or cp.async) and
// turned into the correlation tile smem -> smem (the D x H x W volume never exists in HBM).
// XO: column of the first needed input inside a thread's row window.  TMA boxes must start on a
// 16-byte boundary in W, so for pad-1 kernels the brick origin is 3 columns left of the first tap
// (XO=3: one scalar LDS + aligned LDS.128s); XO=0 otherwise.
// NV: output voxels per thread along W.  4 for the throughput-bound layers (vector LDS, weights
// amortised over 4 voxels); 1 for the small, latency-bound ones (4x the threads, 1/4 of the serial
// FFMA2 chain per thread).
template <int KW, int S, int COG, int CK, bool GWC, bool TMA, int XO, int NV = 4>
__global__ void __launch_bounds__(256, 2) conv_kernel(const __grid_constant__ ConvK p, const __grid_constant__ ConvMaps maps) {
  // TMA destinations must be 128-byte aligned.  Declared aligned (not realigned by pointer
  // arithmetic) so that every derived pointer stays in the .shared state space: LDS, not generic LD.
  extern __shared__ __align__(1024) float smem[];
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
  const int in_elems = (CK * chan_stride + 31) & ~31;   // every buffer is a multiple of 128 bytes
  const int w_elems = (taps * CK * COP + 31) & ~31;
  // smem carve-up: [in0][in1][w0][w1][w2 (GWC)][L staging][R staging][mbarriers]
  const int NS = (TMA && !GWC) ? p.nstages : 2;  // ring depth
  float* s_in0 = smem;
  float* s_w0 = smem + NS * in_elems;
  const int IWR = p.IWR;                  // GWC: right staging row pitch
  float* s_L = s_w0 + (GWC ? 3 : NS) * w_elems;             // GWC only
  const int IWL = p.IWL;                  // GWC: left staging row pitch
  const int l_elems = (CK * p.cpg * IH * IWL + 31) & ~31;
  const int r_elems = (CK * p.cpg * IH * IWR + 31) & ~31;
  float* s_R = s_L + l_elems;                               // GWC only: [CK*cpg][IH][IWR]
  uint64_t* bars = reinterpret_cast<uint64_t*>(GWC ? (s_R + r_elems) : s_L);  // [0..3]: stages; [4]: GWC rows

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

  // GWC staging geometry.  Left rows: [CK*cpg][IH][IWL], column LO <-> x = iw0.  Right rows:
  // [CK*cpg][IH][IWR], column ro <-> x = iw0 - id0 - (ID-1) (the smallest right-image column the
  // brick's disparity range touches).  With TMA both boxes start on a multiple of 4 columns.
  constexpr int LO = TMA ? 3 : 0;
  auto right_origin = [&](int iw0, int id0, int* ro) {
    const int rw0 = iw0 - id0 - (ID - 1);
    const int a = TMA ? (rw0 & ~3) : rw0;  // floor to a multiple of 4 (two's complement: works for negatives)
    *ro = rw0 - a;
    return a;
  };
  auto load_lr = [&](int item) {  // cp.async staging of the left / right feature rows
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
    const int lplane = IH * IWL;
    for (int i = tid; i < nchan * lplane; i += NT) {
      const int c = i / lplane;
      const int rem = i - c * lplane;
      const int hy = rem / IWL;
      const int col = rem - hy * IWL;
      const int h = ih0 + hy, x = iw0 - LO + col;
      const bool ok = (fc0 + c < Cfeat) && h >= 0 && h < p.Hin && x >= 0 && x < p.Win;
      cp_async_4(s_L + i, Lb + (ok ? (long long)(fc0 + c) * sC + (long long)h * sH + x : 0), ok);
    }
    int ro;
    const int rx0 = right_origin(iw0, id0, &ro);
    const int rplane = IH * IWR;
    for (int i = tid; i < nchan * rplane; i += NT) {
      const int c = i / rplane;
      const int rem = i - c * rplane;
      const int hy = rem / IWR;
      const int col = rem - hy * IWR;
      const int h = ih0 + hy, x = rx0 + col;
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
    int ro;
    right_origin(iw0, id0, &ro);
    const int cpg = p.cpg;
    const float inv = 1.0f / (float)cpg;
    const bool pow2 = (cpg & (cpg - 1)) == 0;
    const int lplane = IH * IWL, rplane = IH * IWR;
    // a thread owns (group, row, column) positions and walks the ID disparity planes of each
    for (int i = tid; i < CK * plane; i += NT) {
      const int g = i / plane;
      const int rem = i - g * plane;
      const int hy = rem / IWP;
      const int col = rem - hy * IWP;
      const int h = ih0 + hy, x = iw0 + col;
      const bool ok = (c0 + g < p.Cin) && h >= 0 && h < p.Hin && x >= 0 && x < p.Win;
      float m = 1.f;
      if (mb && ok) m = __ldg(mb + (long long)(c0 + g) * p.imC + (long long)h * p.imH + x);
      const float* lp = s_L + (g * cpg) * lplane + hy * IWL + col + LO;
      const float* rp = s_R + (g * cpg) * rplane + hy * IWR + col + (ID - 1) + ro;
      float* vp = dst + (g * ID) * plane + rem;
      if (cpg == 2) {
        const float l0 = lp[0], l1 = lp[lplane];
        for (int dz = 0; dz < ID; ++dz) {
          const int d = id0 + dz;
          // un-contracted arithmetic: (fea1*fea2).mean(2), submodule.py:147
          const float sum = __fadd_rn(__fmul_rn(l0, rp[-dz]), __fmul_rn(l1, rp[rplane - dz]));
          const bool valid = ok && d >= 0 && d < p.Din && x >= d;
          vp[dz * plane] = valid ? __fmul_rn(__fmul_rn(sum, 0.5f), m) : 0.f;
        }
      } else {
        for (int dz = 0; dz < ID; ++dz) {
          const int d = id0 + dz;
          float v = 0.f;
          if (ok && d >= 0 && d < p.Din && x >= d) {
            float sum = 0.f;
            for (int q = 0; q < cpg; ++q) sum = __fadd_rn(sum, __fmul_rn(lp[q * lplane], rp[q * rplane - dz]));
            v = (pow2 ? sum * inv : sum / (float)cpg) * m;
          }
          vp[dz * plane] = v;
        }
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

  // ---- TMA producers (called by thread 0 only) ----
  auto tma_issue = [&](int item, int stage) {  // ESM_SRC_TENSORS: brick + halo and weight slab of one item
    const int w = blockIdx.x + (item / nch) * gridDim.x;
    const int c0 = (item % nch) * CK;
    const TileCtx t = decode_work(p, w);
    uint64_t* bar = &bars[stage];
    mbar_expect_tx(bar, (unsigned)((CK * chan_stride + taps * CK * COP) * sizeof(float)));
    int rel = c0, k = 0;
    while (k < p.nsrc - 1 && rel >= p.src[k].C) {  // host guarantees chunks never straddle two sources
      rel -= p.src[k].C;
      ++k;
    }
    tma_load_5d(s_in0 + stage * in_elems, &maps.src[k], t.tileW * TW * S - t.pw - XO, t.tileH * p.TH * S - t.ph,
                t.tileD * p.TD * S - t.pd, rel, t.b, bar);
    tma_load_3d(s_w0 + stage * w_elems, &maps.w, t.co_base, c0, t.z * taps, bar);
  };
  auto tma_issue_lr = [&](int item) {  // ESM_SRC_GWC: left / right feature rows + weight slab
    const int w = blockIdx.x + (item / nch) * gridDim.x;
    const int c0 = (item % nch) * CK;
    const TileCtx t = decode_work(p, w);
    const int iw0 = t.tileW * TW * S - t.pw;
    const int ih0 = t.tileH * p.TH * S - t.ph;
    const int id0 = t.tileD * p.TD * S - t.pd;
    uint64_t* bar = &bars[4];
    mbar_expect_tx(bar, (unsigned)((CK * p.cpg * IH * (IWL + IWR) + taps * CK * COP) * sizeof(float)));
    int ro;
    const int rx0 = right_origin(iw0, id0, &ro);
    tma_load_5d(s_L, &maps.src[0], iw0 - LO, ih0, 0, c0 * p.cpg, t.b, bar);
    tma_load_5d(s_R, &maps.src[1], rx0, ih0, 0, c0 * p.cpg, t.b, bar);
    tma_load_3d(s_w0 + (item % 3) * w_elems, &maps.w, t.co_base, c0, t.z * taps, bar);
  };
  // Programmatic dependent launch (common.cuh): no global-memory access before pdl_wait().
  pdl_launch_dependents();
  if (TMA) {
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
  }
  pdl_wait();

  float2 acc[NV][COG / 2];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < COG / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);

  // ---------------- prologue ----------------
  if (n_items > 0) {
    if (GWC) {
      if (TMA) {
        if (tid == 0) tma_issue_lr(0);
        mbar_wait(&bars[4], 0);
      } else {
        load_weights(0, s_w0);
        load_lr(0);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
      }
      build_volume(0, s_in0);
      __syncthreads();
      if (n_items > 1) {
        if (TMA) {
          if (tid == 0) tma_issue_lr(1);
        } else {
          load_weights(1, s_w0 + w_elems);
          load_lr(1);
        }
      }
      if (!TMA) cp_async_commit();
    } else if (TMA) {
      if (tid == 0)
        for (int i = 0; i < NS && i < n_items; ++i) tma_issue(i, i);
    } else {
      load_weights(0, s_w0);
      load_inputs(0, s_in0);
      cp_async_commit();
    }
  }

  for (int item = 0; item < n_items; ++item) {
    const int stage = (TMA && !GWC) ? item % NS : (item & 1);
    const float* s_in = s_in0 + stage * in_elems;
    const float* s_w;
    if (GWC) {
      s_w = s_w0 + (item % 3) * w_elems;
      // rows + weights of item+1 have landed; every warp is done reading V[(item+1)&1] (FFMA2 of item-1,
      // fenced by the barrier that closes each iteration)
      if (TMA) {
        if (item + 1 < n_items) mbar_wait(&bars[4], (item + 1) & 1);
      } else {
        cp_async_wait<0>();
        __syncthreads();
      }
      if (item + 1 < n_items) build_volume(item + 1, s_in0 + ((item + 1) & 1) * in_elems);
      __syncthreads();     // staging rows are free again, V[(item+1)&1] is visible
      if (item + 2 < n_items) {
        if (TMA) {
          if (tid == 0) tma_issue_lr(item + 2);
        } else {
          load_weights(item + 2, s_w0 + ((item + 2) % 3) * w_elems);
          load_lr(item + 2);
        }
      }
      if (!TMA) cp_async_commit();
    } else if (TMA) {
      s_w = s_w0 + stage * w_elems;
      mbar_wait(&bars[stage], (item / NS) & 1);
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
            if (NV == 1) {  // scalar window
#pragma unroll
              for (int q = 0; q < XN; ++q) x[q] = xr[c * chan_stride + XO + q];
            } else if (XO == 0) {
#pragma unroll
              for (int q = 0; q < XL / 4; ++q) {
                const float4 t4 = *reinterpret_cast<const float4*>(xr + c * chan_stride + q * 4);
                x[q * 4 + 0] = t4.x;
                x[q * 4 + 1] = t4.y;
                x[q * 4 + 2] = t4.z;
                x[q * 4 + 3] = t4.w;
              }
            } else {  // window starts at column 3: one scalar, then aligned vectors
              x[0] = xr[c * chan_stride + 3];
#pragma unroll
              for (int q = 0; q < (XN + 2) / 4; ++q) {
                const float4 t4 = *reinterpret_cast<const float4*>(xr + c * chan_stride + 4 + q * 4);
                if (1 + q * 4 + 0 < XL) x[1 + q * 4 + 0] = t4.x;
                if (1 + q * 4 + 1 < XL) x[1 + q * 4 + 1] = t4.y;
                if (1 + q * 4 + 2 < XL) x[1 + q * 4 + 2] = t4.z;
                if (1 + q * 4 + 3 < XL) x[1 + q * 4 + 3] = t4.w;
              }
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
    // Code size matters here: a fully inlined epilogue (7-way activation switch with erff/expf,
    // twice, per output) is ~15k SASS instructions and thrashes the instruction cache once per tile
    // (ncu: 40% stall_no_inst).  The activation is an out-of-line call on 4 values at a time (ILP 4
    // through the erf polynomial), everything else stays in registers.
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
        const int act = p.act, act2 = p.act2;
        const bool post = p.out_mul || p.residual || act2 != ESM_ACT_NONE || p.out_scale != 1.0f;
        if (NV == 1) {
          // one voxel x COG channels per thread: the activation call takes 4 channels at a time
          const int ow = jw0 * osw + t.pz_w;
          if (ow < p.OW) {
#pragma unroll
            for (int j4 = 0; j4 < COG; j4 += 4) {
              const int co0 = t.co_base + cog * COG + j4;
              float rv[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int co = co0 + q;
                const float a = (q & 1) ? acc[0][(j4 + q) / 2].y : acc[0][(j4 + q) / 2].x;
                const float sc = (p.scale && co < p.Cout) ? __ldg(p.scale + co) : 1.f;
                const float sh = (p.shift && co < p.Cout) ? __ldg(p.shift + co) : 0.f;
                rv[q] = fmaf(a, sc, sh);
              }
              float4 r = make_float4(rv[0], rv[1], rv[2], rv[3]);
              if (act != ESM_ACT_NONE) r = apply_act4(r, act);
              rv[0] = r.x; rv[1] = r.y; rv[2] = r.z; rv[3] = r.w;
              if (p.ps == 0) {
                if (post) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const int co = co0 + q;
                    if (co < p.Cout) {
                      if (p.out_mul) rv[q] *= __ldg(p.out_mul + (long long)b * p.omB + (long long)co * p.omC + (long long)oh * p.omH + ow);
                      if (p.residual)
                        rv[q] += __ldg(p.residual + (long long)b * p.oB + (long long)co * p.oC + (long long)od * p.oD + (long long)oh * p.oH + ow);
                    }
                  }
                  if (act2 != ESM_ACT_NONE) {
                    r = apply_act4(make_float4(rv[0], rv[1], rv[2], rv[3]), act2);
                    rv[0] = r.x; rv[1] = r.y; rv[2] = r.z; rv[3] = r.w;
                  }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int co = co0 + q;
                  if (co < p.Cout)
                    p.out[(long long)b * p.oB + (long long)co * p.oC + (long long)od * p.oD + (long long)oh * p.oH + ow] = rv[q] * p.out_scale;
                }
              } else {
                if (act2 != ESM_ACT_NONE) {
                  r = apply_act4(make_float4(rv[0], rv[1], rv[2], rv[3]), act2);
                  rv[0] = r.x; rv[1] = r.y; rv[2] = r.z; rv[3] = r.w;
                }
                const int rr = p.ps;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int co = co0 + q;
                  if (co < p.Cout) {
                    const int c = co / (rr * rr), a = (co / rr) % rr, bb = co % rr;
                    if (p.residual)  // low-resolution residual [B, Cout / r^2, OH, OW], bilinearly upsampled by r
                      rv[q] = __fadd_rn(bilinear_up(p.residual + ((long long)b * (p.Cout / (rr * rr)) + c) * p.OH * p.OW, p.OH, p.OW, oh * rr + a,
                                                    ow * rr + bb, 1.0f / (float)rr), rv[q]);
                    p.out[(long long)b * p.oB + (long long)c * p.oC + (long long)(oh * rr + a) * p.oH + ow * rr + bb] = rv[q] * p.out_scale;
                  }
                }
              }
            }
          }
        } else {
        float4 ps_hold = make_float4(0.f, 0.f, 0.f, 0.f);  // PixelShuffle(2): the even channel of a pair waits for the odd one
#pragma unroll
        for (int j = 0; j < COG; ++j) {
          const int co = t.co_base + cog * COG + j;
          if (co < p.Cout) {
            const float sc = p.scale ? __ldg(p.scale + co) : 1.f;
            const float sh = p.shift ? __ldg(p.shift + co) : 0.f;
            float4 r;
            constexpr int V1 = NV > 1 ? 1 : 0, V2 = NV > 2 ? 2 : 0, V3 = NV > 3 ? 3 : 0;  // (NV == 1 never gets here)
            r.x = fmaf((j & 1) ? acc[0][j / 2].y : acc[0][j / 2].x, sc, sh);
            r.y = fmaf((j & 1) ? acc[V1][j / 2].y : acc[V1][j / 2].x, sc, sh);
            r.z = fmaf((j & 1) ? acc[V2][j / 2].y : acc[V2][j / 2].x, sc, sh);
            r.w = fmaf((j & 1) ? acc[V3][j / 2].y : acc[V3][j / 2].x, sc, sh);
            if (act != ESM_ACT_NONE) r = apply_act4(r, act);
            if (p.ps == 0) {
              const long long obase = (long long)b * p.oB + (long long)co * p.oC + (long long)od * p.oD + (long long)oh * p.oH;
              const int ow0 = jw0 * osw + t.pz_w;
              if (post) {
                const float* om = p.out_mul ? p.out_mul + (long long)b * p.omB + (long long)co * p.omC + (long long)oh * p.omH : nullptr;
                float* rv = &r.x;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                  const int ow = ow0 + v * osw;
                  if (ow < p.OW) {
                    if (om) rv[v] *= __ldg(om + ow);
                    if (p.residual) rv[v] += __ldg(p.residual + obase + ow);
                  }
                }
                if (act2 != ESM_ACT_NONE) r = apply_act4(r, act2);
                r.x *= p.out_scale;
                r.y *= p.out_scale;
                r.z *= p.out_scale;
                r.w *= p.out_scale;
              }
              float* o = p.out + obase;
              if (osw == 1 && ow0 + NV <= p.OW && ((reinterpret_cast<uintptr_t>(o + ow0) & 15) == 0)) {
                *reinterpret_cast<float4*>(o + ow0) = r;
              } else {
                const float* rv = &r.x;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                  const int ow = ow0 + v * osw;
                  if (ow < p.OW) o[ow] = rv[v];
                }
              }
            } else {
              // PixelShuffle(r): channel co -> (c, a, bb); out[c, oh*r + a, ow*r + bb]   (2D only)
              const int rr = p.ps;
              const int c = co / (rr * rr);
              const int a = (co / rr) % rr;
              const int bb = co % rr;
              float* o = p.out + (long long)b * p.oB + (long long)c * p.oC + (long long)(oh * rr + a) * p.oH;
              if (act2 != ESM_ACT_NONE) r = apply_act4(r, act2);
              if (p.residual) {  // low-resolution residual [B, Cout / r^2, OH, OW], bilinearly upsampled by r
                const float* pb = p.residual + ((long long)b * (p.Cout / (rr * rr)) + c) * p.OH * p.OW;
                const float rs = 1.0f / (float)rr;
                float* rw = &r.x;
#pragma unroll
                for (int v = 0; v < NV; ++v)
                  if (jw0 + v < p.OW) rw[v] = __fadd_rn(bilinear_up(pb, p.OH, p.OW, oh * rr + a, (jw0 + v) * rr + bb, rs), rw[v]);
              }
              const float* rv = &r.x;
              // r == 2: channels co (even) and co + 1 are horizontally adjacent output pixels, so the 4 voxels x 2
              // channels of a thread are 8 consecutive floats of one output row: two 16-byte stores instead of eight
              // 4-byte stores 8 bytes apart (the scattered form made the 16->64 layer store-bound: 66 us for 38 MB)
              const bool pair = rr == 2 && NV == 4 && jw0 + NV <= p.OW && ((reinterpret_cast<uintptr_t>(o + jw0 * 2) & 15) == 0);
              if (pair && (j & 1) == 0) {
                ps_hold = r;
              } else if (pair) {
                const float s = p.out_scale;
                float4* o4 = reinterpret_cast<float4*>(o + jw0 * 2);
                o4[0] = make_float4(ps_hold.x * s, r.x * s, ps_hold.y * s, r.y * s);
                o4[1] = make_float4(ps_hold.z * s, r.z * s, ps_hold.w * s, r.w * s);
              } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                  const int ow = jw0 + v;
                  if (ow < p.OW) o[ow * rr + bb] = rv[v] * p.out_scale;
                }
              }
            }
          }
        }
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < COG / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);
    }
    __syncthreads();  // stage (item&1) / V[item&1] may be overwritten from here on
    if (TMA && !GWC && tid == 0 && item + NS < n_items) tma_issue(item + NS, stage);
  }
  if (!TMA) cp_async_wait<0>();
}


typedef void (*conv_fn_t)(const ConvK, const ConvMaps);

// One translation unit per (KW, S) keeps the build parallel; each exports its instantiations.
template <int KW, int S, bool TMA, int XO>
static conv_fn_t pick_cog_ck(int COG, int CK, int nv) {
  if (nv == 1) {  // small-layer variant: instantiated for 8-channel chunks only
    if (CK != 8) return nullptr;
    return COG == 8 ? (conv_fn_t)conv_kernel<KW, S, 8, 8, false, TMA, XO, 1> : (conv_fn_t)conv_kernel<KW, S, 4, 8, false, TMA, XO, 1>;
  }
  if (COG == 8 && CK == 8) return conv_kernel<KW, S, 8, 8, false, TMA, XO>;
  if (COG == 8 && CK == 4) return conv_kernel<KW, S, 8, 4, false, TMA, XO>;
  if (COG == 8 && CK == 1) return conv_kernel<KW, S, 8, 1, false, TMA, XO>;
  if (COG == 4 && CK == 8) return conv_kernel<KW, S, 4, 8, false, TMA, XO>;
  return nullptr;
}
// fused group-wise-correlation input (k3 s1 only); the correlation tile is written by software with
// the aligned (XO=0) layout
template <bool TMA>
static conv_fn_t pick_gwc(int COG, int CK) {
  if (COG != 8) return nullptr;
  return CK == 8 ? (conv_fn_t)conv_kernel<3, 1, 8, 8, true, TMA, 0> : CK == 4 ? (conv_fn_t)conv_kernel<3, 1, 8, 4, true, TMA, 0> : nullptr;
}

// xo: 0 or 3 (only meaningful with tma)
conv_fn_t conv_kernels_k1(int COG, int CK, bool gwc, bool tma, int xo, int nv);
conv_fn_t conv_kernels_k2(int COG, int CK, bool gwc, bool tma, int xo, int nv);
conv_fn_t conv_kernels_k3(int COG, int CK, bool gwc, bool tma, int xo, int nv);
conv_fn_t conv_kernels_k3s2(int COG, int CK, bool gwc, bool tma, int xo, int nv);
conv_fn_t conv_kernels_k5(int COG, int CK, bool gwc, bool tma, int xo, int nv);

}  // namespace esm
