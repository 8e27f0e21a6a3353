// Flat tensor-core path of the convolution family ("tcf"): implicit GEMM on tcgen05.mma (kind::tf32, accumulators in
// TMEM, fp32-grade split-TF32) whose operands arrive by TMA bulk copies ONLY -- no CUDA-core producer warps.
// Replaces BasicConv (submodule.py:12-38) and the bare convs of the hourglass / up_refinement / upsampler / feature
// side (ESMStereo.py:79-125,129-182,185-318) for k1 / k3 stride-1, k3 stride-2 and ConvTranspose k4 s2 p1 layers with
// at least 8 input channels.
//
// What makes that possible is the activation layout between such layers, "PF" (padded-flat, pre-split):
//
//   [B][hi | lo][C/4 quads][Dp * Hp * P positions][4 channels]        fp32, Dp = D + 2 (3D), Hp = H + 2, P = W + 2
//
// * padded: the zero border of the convolution is stored, so tap (kd, kh, kw) of output position f is input position
//   f + (kd-1) Hp P + (kh-1) P + (kw-1) of the SAME flat index space: the im2col rows of 128 consecutive output
//   positions are 128 consecutive input positions, and any window of them is one contiguous 16-byte-per-position
//   range of a channel quad -- a 1D `cp.async.bulk`, no tensor map, no gather;
// * 4-channel quads: a quad plane [position][4] is exactly the K-major, no-swizzle UMMA core-matrix layout (rows 16
//   bytes apart), so the copy lands ready for the tensor core, and a window loaded once serves every kw (and kh) tap
//   by moving the descriptor's start address one row (16 bytes) at a time;
// * pre-split: the producing epilogue stores hi = rna_tf32(x) and lo = x - hi, so the consumer needs no conversion.
//
// Per (tap, 8-channel group) the split scheme costs two MMAs: A_hi x [B_hi | B_lo] (N = 2 NT: main and correction
// columns side by side) and A_lo x B_hi (N = NT) into the correction columns.
//
// A CTA (persistent, 320 threads: 8 epilogue warps, 1 MMA issuer, 1 copy issuer) owns R 128-position tiles per work
// item, either consecutive ("contiguous": one window of 128 R + 2 P + 2 positions feeds all 9 (kh, kw) taps; small
// images) or P apart ("band": R + 2 row windows of 130 positions; wide images), so that an input position is copied
// from L2 ~once per (kd, channel group) instead of once per tap.  Border positions are computed like any other and
// stored as zeros, which is what keeps the output a valid PF tensor.
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace esm {

struct TcfSrc {
  const float* base;    // position 0 of (b = 0, hi, quad 0)
  int ncg;              // 8-channel groups
  long long sB, sHL, sQ;  // strides in floats: batch item, hi -> lo, quad plane
};

struct TcfK {
  TcfSrc src[3];
  int nsrc, ncg;
  int B, Dp, Hp, P, NP, PH;
  int KD, K;            // taps: KD x K x K
  int oz, oy, ox;       // input offset of tap (0, 0, 0) relative to the output position (phase bits are added for transposed)
  int nphase;           // 1; transposed: 4 (2D) / 8 (3D) sub-pixel phases of 2-tap kernels
  int mode;             // 0 contiguous, 1 band
  int R, NPART, NT, ncot, nstages;
  int L, NI, rows_total;  // window length (positions), windows per stage, A rows per (hi|lo, K-half) plane
  int ar, bk;           // A row of (tile r, tap kh, kw) = r * ar + kh * bk + kw
  int items_per_b, J, nbands;
  int total_items;
  const float* w;       // [phase][cot][cg][kd][kh * K + kw][K-half][2 NT rows: hi, lo][4]
  const float* scale;
  const float* shift;
  int act, act2;
  float out_scale, debias;
  int d0, d1, y0, y1, x0, x1;  // valid box of the compute geometry (padded coordinates)
  int omul, osub;       // output coordinate = ((c - c0) * omul + phase bit) / osub  (must divide exactly)
  int oD, oH, oW;       // logical output extent (bounds of the mapped coordinate)
  // PF output (optional)
  float* opf;
  long long o_sB, o_sHL, o_sQ;
  int oHp, oP, od0, oy0, ox0, oCq;
  int same_geom;        // output PF shares the compute geometry: every position is stored (zeros outside the valid box)
  // NCHW output (optional)
  float* out;
  long long oB, oC, oDs, oHs;
  const float* residual;  // NCHW, output strides
  const float* res_pf;    // PF, layout of opf
  int Cout, ps;
};

constexpr int TF_NEW = 16;                  // epilogue warps: four per TMEM lane quadrant
constexpr int TF_MMA_WARP = TF_NEW;         // MMA issuer
constexpr int TF_CP_WARP = TF_NEW + 1;      // first bulk-copy issuer
constexpr int TF_NCP = 2;                   // bulk-copy issuers: warp c takes the planes hl = c (hi / lo), warp 0 also the weights
constexpr int TF_THREADS = 32 * (TF_NEW + 1 + TF_NCP);
constexpr int TF_ACC_COLS = 256;            // TMEM columns per accumulator buffer (two buffers)

struct TfItem {
  int b, phase, cot, fbase, S, mlimit, rlive;
};
__device__ __forceinline__ TfItem tf_decode(const TcfK& p, int item) {
  TfItem t;
  t.cot = item % p.ncot;
  int r = item / p.ncot;
  const int i = r % p.items_per_b;
  r /= p.items_per_b;
  t.phase = r % p.nphase;
  t.b = r / p.nphase;
  if (p.mode == 0) {
    t.fbase = i * (128 * p.R);
    t.S = 128;
    t.mlimit = 128;
    t.rlive = p.R;
  } else {
    const int j = i % p.J;
    const int t2 = i / p.J;
    const int yb = t2 % p.nbands, dz = t2 / p.nbands;
    t.fbase = dz * p.PH + yb * p.R * p.P + 128 * j;
    t.S = p.P;
    t.mlimit = min(128, p.P - 128 * j);
    t.rlive = min(p.R, p.Hp - yb * p.R);
  }
  return t;
}

// n / d for 0 <= n < 2^22 with inv = 1.0f / d (exact after one fix-up in each direction)
__device__ __forceinline__ int tf_div(int n, int d, float inv) {
  int q = (int)((float)n * inv);
  if (q * d > n) --q;
  if ((q + 1) * d <= n) ++q;
  return q;
}

__global__ void __launch_bounds__(TF_THREADS, 1) tcf_conv_kernel(const __grid_constant__ TcfK p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef TC_PROFILE
  const long long tc_t0 = clock64();
#endif
  const int NT = p.NT, NS = p.nstages, KK = p.K * p.K;
  const uint32_t A_PLANE = (uint32_t)p.rows_total * 16;
  const uint32_t A_BYTES = 4 * A_PLANE;
  const uint32_t B_TAP = (uint32_t)NT * 64;
  const uint32_t B_BYTES = (uint32_t)KK * B_TAP;
  const uint32_t STAGE = (A_BYTES + B_BYTES + 127u) & ~127u;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * STAGE);
  uint64_t* empty = full + NS;
  uint64_t* accf = empty + NS;
  uint64_t* acce = accf + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acce + 2);
  float* s_aff = reinterpret_cast<float*>(tmem_slot + 2);  // [2][ncot * NT] scale, shift
  const int nch = p.ncot * NT;
  const int SPI = p.ncg * p.KD;  // ring stages per item

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      tc_mbar_init(&full[i], TF_NCP);
      tc_mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc_mbar_init(&accf[i], 1);
      tc_mbar_init(&acce[i], TF_NEW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * nch; i += TF_THREADS) {
    const int c = i % nch;
    const float* srcp = i < nch ? p.scale : p.shift;
    s_aff[i] = (srcp && c < p.Cout) ? __ldg(srcp + c) : (i < nch ? 1.f : 0.f);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp >= TF_CP_WARP) {
    // ============================ bulk-copy issuers ============================
    // warp hl copies the (hi | lo) planes of every window: 2 NI copies per stage, one per lane; warp 0 also the weights
    const int hl = warp - TF_CP_WARP;
    uint32_t st = 0, ph = 0;
    const int ncopies = 2 * p.NI;
    const uint32_t my_bytes = 2 * A_PLANE + (hl == 0 ? B_BYTES : 0u);
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const TfItem ti = tf_decode(p, item);
      const int pw = ti.phase & 1, phh = (ti.phase >> 1) & 1, pd = (ti.phase >> 2) & 1;
      const long long wbase = (long long)ti.fbase + (long long)(p.oz + pd) * p.PH + (long long)(p.oy + phh) * p.P + (p.ox + pw);
      const float* wsrc = p.w + ((long long)(ti.phase * p.ncot + ti.cot) * SPI) * (KK * NT * 16);
      int s_k = 0, cgl = 0;  // source cursor
      for (int s = 0; s < SPI; ++s) {
        const int kd = s % p.KD;
        if (s > 0 && kd == 0) {
          if (++cgl >= p.src[s_k].ncg) {
            cgl = 0;
            ++s_k;
          }
        }
        tc_mbar_wait(&empty[st], ph ^ 1, 200 + (int)st);
        uint8_t* sa = smem + (size_t)st * STAGE;
        if (lane == 0) tc_mbar_expect_tx(&full[st], my_bytes);
        __syncwarp();
        const TcfSrc& sv = s_k == 0 ? p.src[0] : (s_k == 1 ? p.src[1] : p.src[2]);
        const float* sb = sv.base + (long long)ti.b * sv.sB + (long long)hl * sv.sHL + (long long)(cgl * 2) * sv.sQ + (wbase + (long long)kd * p.PH) * 4;
        for (int c = lane; c < ncopies; c += 32) {
          const int i = c >> 1, kh_ = c & 1;  // window, K-half
          const float* g = sb + (long long)kh_ * sv.sQ + (long long)i * p.P * 4;
          tc_bulk_g2s(sa + (size_t)(hl * 2 + kh_) * A_PLANE + (size_t)i * p.L * 16, g, (uint32_t)p.L * 16, &full[st]);
        }
        if (hl == 0 && lane == 31) tc_bulk_g2s(sa + A_BYTES, wsrc + (long long)s * (KK * NT * 16), B_BYTES, &full[st]);
        __syncwarp();
        if (++st == (uint32_t)NS) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == TF_MMA_WARP) {
    // ============================ MMA issuer ============================
    // One elected thread runs the whole loop (waits, MMAs, commits): under a branch that ptxas can tie to elect.sync
    // the descriptor arithmetic stays in the uniform datapath; `if (leader)` around each MMA group inside a
    // warp-wide loop cost ~8 R2URs per MMA (65 clk per MMA issued against 46 executed).
    if (tc_elect()) {
      const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * NT) >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t a0 = tc_desc(tc_smem_u32(smem), A_PLANE, 128);
      const uint64_t b0 = tc_desc(tc_smem_u32(smem) + A_BYTES, (uint32_t)NT * 32, 128);
      const uint32_t lo_off = (2 * A_PLANE) >> 4;
      const uint32_t bt16 = B_TAP >> 4;
      const int K = p.K, R = p.R, ar = p.ar, bk = p.bk, NPART = p.NPART;
      uint32_t st = 0, ph = 0, ai = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const uint32_t ab = ai & 1, aph = (ai >> 1) & 1;
        tc_mbar_wait(&acce[ab], aph ^ 1, 400 + (int)ab);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int part = 0;
        for (int s = 0; s < SPI; ++s) {
          tc_mbar_wait(&full[st], ph, 500 + (int)st);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t a_st = a0 + (uint64_t)((st * STAGE) >> 4);
          const uint64_t b_st = b0 + (uint64_t)((st * STAGE) >> 4);
          const uint32_t first = s < NPART ? 0u : 1u;
          const uint32_t d0 = tmem + ab * TF_ACC_COLS + (uint32_t)(part * 2 * NT);
          for (int r = 0; r < R; ++r) {
            const uint32_t d = d0 + (uint32_t)(r * NPART * 2 * NT);
            uint64_t a_kh = a_st + (uint64_t)(r * ar);
            uint64_t b_t = b_st;
            for (int kh = 0; kh < K; ++kh) {
              if (K == 3) {
                tc_mma(d, a_kh, b_t, idesc1, kh == 0 ? first : 1u);
                tc_mma(d + NT, a_kh + lo_off, b_t, idesc2, 1u);
                tc_mma(d, a_kh + 1, b_t + bt16, idesc1, 1u);
                tc_mma(d + NT, a_kh + 1 + lo_off, b_t + bt16, idesc2, 1u);
                tc_mma(d, a_kh + 2, b_t + 2 * bt16, idesc1, 1u);
                tc_mma(d + NT, a_kh + 2 + lo_off, b_t + 2 * bt16, idesc2, 1u);
              } else if (K == 2) {
                tc_mma(d, a_kh, b_t, idesc1, kh == 0 ? first : 1u);
                tc_mma(d + NT, a_kh + lo_off, b_t, idesc2, 1u);
                tc_mma(d, a_kh + 1, b_t + bt16, idesc1, 1u);
                tc_mma(d + NT, a_kh + 1 + lo_off, b_t + bt16, idesc2, 1u);
              } else {
                tc_mma(d, a_kh, b_t, idesc1, first);
                tc_mma(d + NT, a_kh + lo_off, b_t, idesc2, 1u);
              }
              a_kh += (uint64_t)bk;
              b_t += (uint64_t)(K * bt16);
            }
          }
          tc_commit(&empty[st]);
          if (++part == NPART) part = 0;
          if (++st == (uint32_t)NS) {
            st = 0;
            ph ^= 1;
          }
        }
        tc_commit(&accf[ab]);
        ++ai;
      }
    }
  } else {
    // ============================ epilogue ============================
    // the four warps w, w + 4, w + 8, w + 12 share TMEM lane quadrant q = w % 4 and take the (tile, 8-channel chunk)
    // units round-robin; a unit is read with two x8 loads per partial accumulator (main + correction columns).
    const int q = warp & 3, sub = warp >> 2;
    const int m = q * 32 + lane;
    const int nchunk = NT >> 3;
    const int units = p.R * nchunk;
    const int act = p.act;
    const float invP = 1.0f / (float)p.P, invH = 1.0f / (float)p.Hp;
    uint32_t ai = 0;
#ifdef TC_PROFILE
    long long prof_busy = 0;
#endif
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const TfItem ti = tf_decode(p, item);
      const uint32_t ab = ai & 1, aph = (ai >> 1) & 1;
      const int pw = ti.phase & 1, phh = (ti.phase >> 1) & 1, pd = (ti.phase >> 2) & 1;
      // position of row m of tile 0 (one pair of integer divisions per item; the other tiles are derived from it)
      const long long f0 = (long long)ti.fbase + m;
      int d_0, y_0, x_0;
      {
        const int fi = (int)(f0 < (long long)p.NP ? f0 : 0);
        d_0 = fi / p.PH;
        const int rem = fi - d_0 * p.PH;
        y_0 = rem / p.P;
        x_0 = rem - y_0 * p.P;
      }
      tc_mbar_wait(&accf[ab], aph, 600 + (int)ab);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef TC_PROFILE
      const long long te0 = clock64();
#endif
      for (int u = sub; u < units; u += 4) {
        const int r = u / nchunk, ck = u - r * nchunk;
        const long long f = f0 + (long long)r * ti.S;
        const bool live = r < ti.rlive && m < ti.mlimit && f < (long long)p.NP;
        int d = d_0, y = y_0, x = x_0;
        if (p.mode == 0) {
          const int t = x_0 + 128 * r;
          const int yq = tf_div(t, p.P, invP);
          x = t - yq * p.P;
          const int t2 = y_0 + yq;
          const int dq = tf_div(t2, p.Hp, invH);
          y = t2 - dq * p.Hp;
          d = d_0 + dq;
        } else {
          y = y_0 + r;
        }
        const bool valid = live && d >= p.d0 && d < p.d1 && y >= p.y0 && y < p.y1 && x >= p.x0 && x < p.x1;
        bool ok = valid;
        int zo = d - p.d0, yo = y - p.y0, xo = x - p.x0;
        if (p.omul != 1 || p.osub != 1 || p.nphase > 1) {
          zo = zo * (p.Dp > 1 ? p.omul : 1) + pd;
          yo = yo * p.omul + phh;
          xo = xo * p.omul + pw;
          if (p.osub == 2) {
            ok = ok && !(yo & 1) && !(xo & 1) && (p.Dp == 1 || !(zo & 1));
            if (p.Dp > 1) zo >>= 1;
            yo >>= 1;
            xo >>= 1;
          }
        }
        ok = ok && zo < p.oD && yo < p.oH && xo < p.oW;
        const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + ab * TF_ACC_COLS + (uint32_t)(r * p.NPART * 2 * NT + ck * 8);
        float mv[8], cv[8];
        tc_ld8(tb, mv);
        tc_ld8(tb + NT, cv);
        tc_ld_wait();
        for (int part = 1; part < p.NPART; ++part) {
          float a[8], c[8];
          tc_ld8(tb + part * 2 * NT, a);
          tc_ld8(tb + part * 2 * NT + NT, c);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            mv[j] += a[j];
            cv[j] += c[j];
          }
        }
        const int c0 = ti.cot * NT + ck * 8;  // first output channel of this chunk
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(fmaf(mv[j], p.debias, cv[j]), s_aff[c0 + j], s_aff[nch + c0 + j]);
        if (act == ESM_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = tc_gelu(v[j]);
        } else if (act == ESM_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = tc_silu(v[j]);
        } else if (act == ESM_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (act != ESM_ACT_NONE) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 t4 = apply_act4(make_float4(v[4 * h2], v[4 * h2 + 1], v[4 * h2 + 2], v[4 * h2 + 3]), act);
            v[4 * h2] = t4.x; v[4 * h2 + 1] = t4.y; v[4 * h2 + 2] = t4.z; v[4 * h2 + 3] = t4.w;
          }
        }
        if (p.ps == 2) {
          // PixelShuffle(2) of a 2D layer: a quad's channels are the 2 x 2 pixels of channel c / 4 at (2y, 2x)
          if (p.act2 == ESM_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = tc_silu(v[j]);
          }
          if (ok) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              if (c0 + 4 * h2 < p.Cout) {
                float* o = p.out + (long long)ti.b * p.oB + (long long)((c0 >> 2) + h2) * p.oC + (long long)(2 * yo) * p.oHs + 2 * xo;
                *reinterpret_cast<float2*>(o) = make_float2(v[4 * h2] * p.out_scale, v[4 * h2 + 1] * p.out_scale);
                *reinterpret_cast<float2*>(o + p.oHs) = make_float2(v[4 * h2 + 2] * p.out_scale, v[4 * h2 + 3] * p.out_scale);
              }
            }
          }
          continue;
        }
        long long opos = 0;  // PF output position (floats) of the chunk's first quad
        if (p.opf) {
          const long long of = p.same_geom ? f : ((long long)(zo + p.od0) * p.oHp + (yo + p.oy0)) * p.oP + (xo + p.ox0);
          opos = (long long)ti.b * p.o_sB + (long long)(c0 >> 2) * p.o_sQ + of * 4;
        }
        const long long npos = (long long)ti.b * p.oB + (long long)c0 * p.oC + (long long)zo * p.oDs + (long long)yo * p.oHs + xo;
        if (ok) {
          if (p.residual) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c0 + j < p.Cout) v[j] += __ldg(p.residual + npos + j * p.oC);
          }
          if (p.res_pf) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              if ((c0 >> 2) + h2 >= p.oCq) break;
              const float4 rh = __ldg(reinterpret_cast<const float4*>(p.res_pf + opos + h2 * p.o_sQ));
              const float4 rl = __ldg(reinterpret_cast<const float4*>(p.res_pf + opos + h2 * p.o_sQ + p.o_sHL));
              v[4 * h2] += rh.x + rl.x; v[4 * h2 + 1] += rh.y + rl.y; v[4 * h2 + 2] += rh.z + rl.z; v[4 * h2 + 3] += rh.w + rl.w;
            }
          }
          if (p.act2 != ESM_ACT_NONE) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const float4 t4 = apply_act4(make_float4(v[4 * h2], v[4 * h2 + 1], v[4 * h2 + 2], v[4 * h2 + 3]), p.act2);
              v[4 * h2] = t4.x; v[4 * h2 + 1] = t4.y; v[4 * h2 + 2] = t4.z; v[4 * h2 + 3] = t4.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= p.out_scale;
        }
        if (p.out && ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c0 + j < p.Cout) p.out[npos + j * p.oC] = v[j];
        }
        if (p.opf && (p.same_geom ? live : ok)) {
          const bool keep = p.same_geom ? (valid && ok) : true;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if ((c0 >> 2) + h2 >= p.oCq) break;
            float hv[4], lv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float x_ = (keep && c0 + 4 * h2 + j < p.Cout) ? v[4 * h2 + j] : 0.f;
              hv[j] = tc_rna(x_);
              lv[j] = x_ - hv[j];
            }
            *reinterpret_cast<float4*>(p.opf + opos + h2 * p.o_sQ) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<float4*>(p.opf + opos + h2 * p.o_sQ + p.o_sHL) = make_float4(lv[0], lv[1], lv[2], lv[3]);
          }
        }
      }
      // every TMEM read of this item is complete (tc_ld_wait above): hand the accumulator buffer back
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&acce[ab]);
      ++ai;
#ifdef TC_PROFILE
      prof_busy += clock64() - te0;
#endif
    }
#ifdef TC_PROFILE
    if (blockIdx.x == 0 && lane == 0) tc_prof_wait[warp][1] = (unsigned long long)prof_busy;
#endif
  }
#ifdef TC_PROFILE
  if (blockIdx.x == 0 && lane == 0) {
    printf("tcf_prof warp %2d: done at %8lld clk; waits: empty %8llu  acce %8llu  full %8llu  accf %8llu | issue %8llu epi busy %8llu\n", warp,
           clock64() - tc_t0, tc_prof_wait[warp][2], tc_prof_wait[warp][4], tc_prof_wait[warp][5], tc_prof_wait[warp][6], tc_prof_wait[warp][0],
           tc_prof_wait[warp][1]);
    for (int i = 0; i < 8; ++i) tc_prof_wait[warp][i] = 0;
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------
// PF <-> NCHW converters, weight packing
// ------------------------------------------------------------------------------------------
struct PfGeom {
  int B, C, Cq, Dp, Hp, P, NP;
  int d0, d1, y0, y1, x0, x1;
};

__global__ void pf_from_nchw_kernel(const float* __restrict__ x, long long sB, long long sC, long long sD, long long sH, float* __restrict__ o,
                                    PfGeom g, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = (int)(i % g.NP);
  long long r = i / g.NP;
  const int qd = (int)(r % g.Cq);
  const int b = (int)(r / g.Cq);
  const int PH = g.Hp * g.P;
  const int d = f / PH, rem = f - d * PH, y = rem / g.P, xx = rem - y * g.P;
  const bool valid = d >= g.d0 && d < g.d1 && y >= g.y0 && y < g.y1 && xx >= g.x0 && xx < g.x1;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const float* s = x + (long long)b * sB + (long long)(d - g.d0) * sD + (long long)(y - g.y0) * sH + (xx - g.x0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (qd * 4 + j < g.C) v[j] = __ldg(s + (long long)(qd * 4 + j) * sC);
  }
  float4 hi, lo;
  hi.x = tc_rna(v[0]); hi.y = tc_rna(v[1]); hi.z = tc_rna(v[2]); hi.w = tc_rna(v[3]);
  lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
  const long long sQ = (long long)g.NP * 4, sHL = sQ * g.Cq;
  float* op = o + (long long)b * 2 * sHL + (long long)qd * sQ + (long long)f * 4;
  *reinterpret_cast<float4*>(op) = hi;
  *reinterpret_cast<float4*>(op + sHL) = lo;
}

__global__ void pf_to_nchw_kernel(const float* __restrict__ pf, PfGeom g, float* __restrict__ out, long long sB, long long sC, long long sD,
                                  long long sH, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W = g.x1 - g.x0, H = g.y1 - g.y0, D = g.d1 - g.d0;
  long long r = i;
  const int x = (int)(r % W);
  r /= W;
  const int y = (int)(r % H);
  r /= H;
  const int d = (int)(r % D);
  r /= D;
  const int c = (int)(r % g.C);
  const int b = (int)(r / g.C);
  const long long sQ = (long long)g.NP * 4, sHL = sQ * g.Cq;
  const long long f = ((long long)(d + g.d0) * g.Hp + (y + g.y0)) * g.P + (x + g.x0);
  const float* ip = pf + (long long)b * 2 * sHL + (long long)(c >> 2) * sQ + f * 4 + (c & 3);
  out[(long long)b * sB + (long long)c * sC + (long long)d * sD + (long long)y * sH + x] = ip[0] + ip[sHL];
}

// packed layout: [phase][cot][cg][kd][kh * K + kw][K-half][2 NT rows: hi then lo][4]
__global__ void pack_tcf_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int kd, int kh, int kw, int transposed,
                                int KD, int K, int nphase, int ncot, int NT, int ncg, int nsrc, int c0, int c1, int c2, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long r = i;
  const int k4 = (int)(r % 4);
  r /= 4;
  const int row = (int)(r % NT);
  r /= NT;
  const int khalf = (int)(r % 2);
  r /= 2;
  const int tap = (int)(r % (K * K));
  r /= K * K;
  const int tdd = (int)(r % KD);
  r /= KD;
  const int cg = (int)(r % ncg);
  r /= ncg;
  const int cot = (int)(r % ncot);
  const int z = (int)(r / ncot);  // phase
  const int thh = tap / K, tw = tap % K;
  // channel gp of the padded concatenation -> channel ci of the weight (or none)
  const int gp = cg * 8 + khalf * 4 + k4;
  const int cs[3] = {c0, c1, c2};
  int ci = -1, pad_base = 0, real_base = 0;
  for (int s = 0; s < nsrc; ++s) {
    const int padded = (cs[s] + 7) / 8 * 8;
    if (gp < pad_base + padded) {
      if (gp - pad_base < cs[s]) ci = real_base + gp - pad_base;
      break;
    }
    pad_base += padded;
    real_base += cs[s];
  }
  const int co = cot * NT + row;
  float v = 0.f;
  if (co < Cout && ci >= 0 && ci < Cin) {
    if (!transposed) {
      v = w[(((long long)(co * Cin + ci) * kd + tdd) * kh + thh) * kw + tw];
    } else {
      const int pzw = z & 1, pzh = (z >> 1) & 1, pzd = (z >> 2) & 1;
      const int kkw = 3 - pzw - 2 * tw;
      const int kkh = 3 - pzh - 2 * thh;
      const int kkd = (kd == 4) ? 3 - pzd - 2 * tdd : 0;
      v = w[(((long long)(ci * Cout + co) * kd + kkd) * kh + kkh) * kw + kkw];  // [Cin, Cout, k, k, k]
    }
  }
  const float hi = tc_rna(v);
  const long long slab = ((((long long)(z * ncot + cot) * ncg + cg) * KD + tdd) * (K * K) + tap) * (16 * NT);
  float* o = out + slab + (long long)khalf * (2 * NT * 4) + (long long)row * 4 + k4;
  o[0] = hi;
  o[NT * 4] = v - hi;
}

static void tcf_weight_geom(int Cout, int nsrc, const int* srcC, int kd, int kh, int transposed, int* KD, int* K, int* nphase, int* ncot, int* NT,
                            int* ncg) {
  const int CoutPad = round_up(Cout, 8);
  *ncot = ceil_div(CoutPad, 64);
  *NT = round_up(ceil_div(CoutPad, *ncot), 8);
  int g = 0;
  for (int i = 0; i < nsrc; ++i) g += ceil_div(srcC[i], 8);
  *ncg = g;
  if (transposed) {
    *KD = kd == 4 ? 2 : 1;
    *K = 2;
    *nphase = kd == 4 ? 8 : 4;
  } else {
    *KD = kd;
    *K = kh;
    *nphase = 1;
  }
}

static long long tcf_launches = 0;

}  // namespace esm

using namespace esm;

extern "C" long long esm_tcf_conv_launches(void) { return tcf_launches; }

extern "C" long long esm_pf_guard_elems(int Dp, int Hp, int P) {
  return 4ll * ((Dp > 1 ? (long long)Hp * P : 0) + 10ll * P + 2048);
}

extern "C" long long esm_pf_elems(int B, int C, int Dp, int Hp, int P) {
  return (long long)B * 2 * (2 * ceil_div(C, 8)) * Dp * Hp * P * 4;
}

static int pf_geom(const esm_pf_t* t, PfGeom* g, const char* what) {
  ESM_REQUIRE(t && t->data, "%s: null PF tensor", what);
  ESM_REQUIRE(t->B > 0 && t->C > 0 && t->Dp > 0 && t->Hp > 2 && t->P > 2, "%s: empty PF tensor", what);
  ESM_REQUIRE(t->d0 >= 0 && t->d1 <= t->Dp && t->y0 >= 1 && t->y1 <= t->Hp - 1 && t->x0 >= 1 && t->x1 <= t->P - 1 && t->d0 < t->d1 &&
                  t->y0 < t->y1 && t->x0 < t->x1 && (t->Dp == 1 || (t->d0 >= 1 && t->d1 <= t->Dp - 1)),
              "%s: valid box must leave a one-position border", what);
  ESM_REQUIRE((long long)t->Dp * t->Hp * t->P < (1ll << 30), "%s: PF plane too large", what);
  ESM_REQUIRE((reinterpret_cast<uintptr_t>(t->data) & 15) == 0, "%s: PF data must be 16-byte aligned", what);
  g->B = t->B;
  g->C = t->C;
  g->Cq = 2 * ceil_div(t->C, 8);
  g->Dp = t->Dp;
  g->Hp = t->Hp;
  g->P = t->P;
  g->NP = t->Dp * t->Hp * t->P;
  g->d0 = t->d0; g->d1 = t->d1; g->y0 = t->y0; g->y1 = t->y1; g->x0 = t->x0; g->x1 = t->x1;
  return ESM_OK;
}

extern "C" int esm_pf_from_nchw_f32(const float* x, long long sB, long long sC, long long sD, long long sH, const esm_pf_t* out, void* stream) {
  PfGeom g;
  if (int e = pf_geom(out, &g, "pf_from_nchw")) return e;
  ESM_REQUIRE(x, "pf_from_nchw: null input");
  const long long total = (long long)g.B * g.Cq * g.NP;
  pf_from_nchw_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(x, sB, sC, sD, sH, out->data, g, total);
  return check_launch("pf_from_nchw");
}

extern "C" int esm_pf_to_nchw_f32(const esm_pf_t* in, float* out, long long sB, long long sC, long long sD, long long sH, void* stream) {
  PfGeom g;
  if (int e = pf_geom(in, &g, "pf_to_nchw")) return e;
  ESM_REQUIRE(out, "pf_to_nchw: null output");
  const long long total = (long long)g.B * g.C * (g.d1 - g.d0) * (g.y1 - g.y0) * (g.x1 - g.x0);
  pf_to_nchw_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(in->data, g, out, sB, sC, sD, sH, total);
  return check_launch("pf_to_nchw");
}

extern "C" long long esm_packed_weight_pf_elems(int Cout, int nsrc, const int* srcC, int kd, int kh, int kw, int transposed) {
  if (nsrc < 1 || nsrc > 3 || !srcC) return 0;
  int KD, K, nphase, ncot, NT, ncg;
  tcf_weight_geom(Cout, nsrc, srcC, kd, kh, transposed, &KD, &K, &nphase, &ncot, &NT, &ncg);
  return (long long)nphase * ncot * ncg * KD * K * K * 16 * NT;
}

extern "C" int esm_pack_conv_weight_pf_f32(const float* w, float* packed, int Cout, int nsrc, const int* srcC, int kd, int kh, int kw,
                                           int transposed, void* stream) {
  ESM_REQUIRE(w && packed && Cout > 0 && nsrc >= 1 && nsrc <= 3 && srcC, "pack_conv_weight_pf: bad arguments");
  ESM_REQUIRE(kh == kw, "pack_conv_weight_pf: square kernels only");
  if (transposed) ESM_REQUIRE((kd == 4 || kd == 1) && kh == 4, "pack_conv_weight_pf: transposed conv must be k4");
  int KD, K, nphase, ncot, NT, ncg, Cin = 0;
  for (int i = 0; i < nsrc; ++i) Cin += srcC[i];
  tcf_weight_geom(Cout, nsrc, srcC, kd, kh, transposed, &KD, &K, &nphase, &ncot, &NT, &ncg);
  const long long total = (long long)nphase * ncot * ncg * KD * K * K * 2 * NT * 4;  // one thread per (hi, lo) pair
  pack_tcf_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(w, packed, Cout, Cin, kd, kh, kw, transposed, KD, K, nphase,
                                                                                      ncot, NT, ncg, nsrc, srcC[0], nsrc > 1 ? srcC[1] : 0,
                                                                                      nsrc > 2 ? srcC[2] : 0, total);
  return check_launch("pack_conv_weight_pf");
}

extern "C" int esm_conv_pf_f32(const esm_conv_pf_t* d, void* stream) {
  ESM_REQUIRE(d && d->weight, "conv_pf: null descriptor / weight");
  ESM_REQUIRE(d->nsrc >= 1 && d->nsrc <= 3, "conv_pf: nsrc must be 1..3");
  ESM_REQUIRE(d->out_pf.data || d->out, "conv_pf: no output");
  PfGeom g0;
  if (int e = pf_geom(&d->src[0], &g0, "conv_pf(src0)")) return e;
  int srcC[3] = {0, 0, 0};
  for (int i = 0; i < d->nsrc; ++i) {
    PfGeom gi;
    if (int e = pf_geom(&d->src[i], &gi, "conv_pf(src)")) return e;
    ESM_REQUIRE(gi.B == g0.B && gi.Dp == g0.Dp && gi.Hp == g0.Hp && gi.P == g0.P, "conv_pf: sources differ in geometry");
    srcC[i] = gi.C;
  }
  const bool is3d = g0.Dp > 1;
  const int tr = d->transposed;
  if (tr) {
    ESM_REQUIRE(d->kh == 4 && d->kw == 4 && (d->kd == (is3d ? 4 : 1)) && d->stride == 2, "conv_pf: transposed conv supports k4 s2 p1 only");
  } else {
    ESM_REQUIRE(d->kh == d->kw && (d->kh == 1 || d->kh == 3) && (d->kd == 1 || (is3d && d->kd == d->kh)), "conv_pf: k1 / k3 kernels only");
    ESM_REQUIRE(d->stride == 1 || (d->stride == 2 && d->kh == 3), "conv_pf: stride 1, or 2 with k3");
  }
  TcfK k;
  memset(&k, 0, sizeof(k));
  tcf_weight_geom(d->Cout, d->nsrc, srcC, d->kd, d->kh, tr, &k.KD, &k.K, &k.nphase, &k.ncot, &k.NT, &k.ncg);
  k.nsrc = d->nsrc;
  for (int i = 0; i < d->nsrc; ++i) {
    k.src[i].base = d->src[i].data;
    k.src[i].ncg = ceil_div(srcC[i], 8);
    k.src[i].sQ = (long long)g0.NP * 4;
    k.src[i].sHL = k.src[i].sQ * 2 * k.src[i].ncg;
    k.src[i].sB = 2 * k.src[i].sHL;
  }
  k.B = g0.B; k.Dp = g0.Dp; k.Hp = g0.Hp; k.P = g0.P; k.NP = g0.NP; k.PH = g0.Hp * g0.P;
  // tap (0,0,0) offset: regular convs are centred (pad = k / 2); a transposed phase p reads offsets p - 1, p
  k.oz = tr ? (k.KD == 2 ? -1 : 0) : -(k.KD / 2);
  k.oy = tr ? -1 : -(k.K / 2);
  k.ox = k.oy;
  // the valid box the layer computes over
  k.d0 = d->d0; k.d1 = d->d1; k.y0 = d->y0; k.y1 = d->y1; k.x0 = d->x0; k.x1 = d->x1;
  ESM_REQUIRE(k.d0 >= 0 && k.d1 <= g0.Dp && k.y0 >= 1 && k.y1 <= g0.Hp - 1 && k.x0 >= 1 && k.x1 <= g0.P - 1 && k.d0 < k.d1 && k.y0 < k.y1 &&
                  k.x0 < k.x1,
              "conv_pf: compute box outside the geometry");
  k.omul = tr ? 2 : 1;
  k.osub = (!tr && d->stride == 2) ? 2 : 1;
  k.oD = d->oD; k.oH = d->oH; k.oW = d->oW;
  ESM_REQUIRE(k.oD > 0 && k.oH > 0 && k.oW > 0, "conv_pf: empty output extent");
  k.w = d->weight;
  k.scale = d->scale;
  k.shift = d->shift;
  k.act = d->act;
  k.act2 = d->act2;
  k.out_scale = d->out_scale;
  k.Cout = d->Cout;
  k.ps = d->pixel_shuffle;
  ESM_REQUIRE(k.ps == 0 || (k.ps == 2 && !is3d && d->out && !d->out_pf.data && !d->residual && !d->res_pf && d->Cout % 4 == 0 && !tr && d->stride == 1),
              "conv_pf: pixel_shuffle needs a 2D stride-1 layer with an NCHW output only");
  if (d->out_pf.data) {
    PfGeom go;
    if (int e = pf_geom(&d->out_pf, &go, "conv_pf(out)")) return e;
    ESM_REQUIRE(go.B == g0.B && go.C == d->Cout, "conv_pf: output PF batch / channels mismatch");
    k.opf = d->out_pf.data;
    k.o_sQ = (long long)go.NP * 4;
    k.o_sHL = k.o_sQ * go.Cq;
    k.o_sB = 2 * k.o_sHL;
    k.oHp = go.Hp; k.oP = go.P; k.od0 = go.d0; k.oy0 = go.y0; k.ox0 = go.x0; k.oCq = go.Cq;
    k.same_geom = (!tr && d->stride == 1 && go.Dp == g0.Dp && go.Hp == g0.Hp && go.P == g0.P && go.d0 == k.d0 && go.y0 == k.y0 && go.x0 == k.x0 &&
                   go.d1 == k.d1 && go.y1 == k.y1 && go.x1 == k.x1)
                      ? 1
                      : 0;
    if (!k.same_geom)
      ESM_REQUIRE(go.d1 - go.d0 == k.oD && go.y1 - go.y0 == k.oH && go.x1 - go.x0 == k.oW, "conv_pf: output PF valid box != output extent");
    // channels beyond Cout up to the quad count are written as zeros only when NT covers them
    ESM_REQUIRE(go.Cq * 4 <= k.ncot * k.NT, "conv_pf: output quads not covered by the channel tiles");
  }
  if (d->res_pf) {
    ESM_REQUIRE(d->out_pf.data, "conv_pf: res_pf needs a PF output (it shares its layout)");
    k.res_pf = d->res_pf;
  }
  k.out = d->out;
  k.oB = d->oB; k.oC = d->oC; k.oDs = d->oDs; k.oHs = d->oHs;
  k.residual = d->residual;
  ESM_REQUIRE(!d->residual || d->out, "conv_pf: NCHW residual needs an NCHW output (it shares its strides)");

  // ---- plan ----
  int num_sms = 0;
  {
    static int sms_cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
      cudaGetLastError();
      set_error("conv_pf: no CUDA device");
      return ESM_ERR_CUDA;
    }
    if (dev < 0 || dev >= 64 || sms_cache[dev] == 0) {
      int n = 0;
      if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("conv_pf: no CUDA device");
        return ESM_ERR_CUDA;
      }
      if (dev >= 0 && dev < 64) sms_cache[dev] = n;
      num_sms = n;
    } else {
      num_sms = sms_cache[dev];
    }
  }
  const int SPI = k.ncg * k.KD, KK = k.K * k.K;
  // partial accumulators: keep a chain of accumulates short (the tensor core truncates its fp32 accumulator)
  int npart = ceil_div(SPI * KK, 48);
  const int maxslots = TF_ACC_COLS / (2 * k.NT);
  ESM_REQUIRE(maxslots >= 1, "conv_pf: channel tile too wide");
  if (npart > maxslots) npart = maxslots;
  if (getenv("ESM_TCF_NPART")) npart = atoi(getenv("ESM_TCF_NPART"));
  if (npart > maxslots) npart = maxslots;
  if (npart > SPI) npart = SPI;
  if (npart < 1) npart = 1;
  k.NPART = npart;
  const int rmax = maxslots / npart;
  const int env_mode = getenv("ESM_TCF_MODE") ? atoi(getenv("ESM_TCF_MODE")) : -1;
  const int env_r = getenv("ESM_TCF_R") ? atoi(getenv("ESM_TCF_R")) : 0;
  k.mode = env_mode >= 0 ? env_mode : (g0.P >= 400 ? 1 : 0);
  const long long units = (long long)g0.B * k.nphase * k.ncot;
  auto nitems = [&](int mode, int R) -> long long {
    if (mode == 0) return ceil_div(ceil_div(g0.NP, 128), R);
    return (long long)g0.Dp * ceil_div(g0.Hp, R) * ceil_div(g0.P, 128);
  };
  auto stage_bytes = [&](int mode, int R) -> size_t {
    const int L = mode == 0 ? 128 * R + (k.K - 1) * g0.P + (k.K - 1) : 128 + k.K - 1;
    const int NI = mode == 0 ? 1 : R + k.K - 1;
    return (((size_t)4 * NI * L * 16 + (size_t)KK * k.NT * 64) + 127) & ~(size_t)127;
  };
  const size_t limit = 220 * 1024;
  int R = 1;
  for (int cand = (rmax > 4 ? 4 : rmax); cand >= 1; --cand) {
    if (2 * stage_bytes(k.mode, cand) > limit) continue;
    // the largest R that still gives every SM ~2 items (or R = 1)
    if (cand == 1 || nitems(k.mode, cand) * units >= 2ll * num_sms) {
      R = cand;
      break;
    }
  }
  if (env_r > 0 && env_r <= rmax) R = env_r;
  ESM_REQUIRE(2 * stage_bytes(k.mode, R) <= limit, "conv_pf: stage does not fit shared memory (P=%d)", g0.P);
  k.R = R;
  if (k.mode == 0) {
    k.L = 128 * R + (k.K - 1) * g0.P + (k.K - 1);
    k.NI = 1;
    k.ar = 128;
    k.bk = g0.P;
    k.items_per_b = ceil_div(ceil_div(g0.NP, 128), R);
    k.J = 1;
    k.nbands = 1;
  } else {
    k.L = 128 + k.K - 1;
    k.NI = R + k.K - 1;
    k.ar = k.L;
    k.bk = k.L;
    k.J = ceil_div(g0.P, 128);
    k.nbands = ceil_div(g0.Hp, R);
    k.items_per_b = g0.Dp * k.nbands * k.J;
  }
  k.rows_total = k.NI * k.L;
  const size_t sb = stage_bytes(k.mode, R);
  int ns = (int)(limit / sb);
  if (ns > 4) ns = 4;
  if (ns > SPI + 1) ns = SPI + 1 > 2 ? SPI + 1 : 2;
  k.nstages = ns;
  const long long total_items = units * k.items_per_b;
  ESM_REQUIRE(total_items < (1ll << 30), "conv_pf: too many work items");
  k.total_items = (int)total_items;
  k.debias = 1.0f + TC_TRUNC_BIAS * (float)(ceil_div(SPI, npart) * KK);
  const size_t smem = (size_t)ns * sb + (2 * ns + 4) * 8 + 16 + (size_t)2 * k.ncot * k.NT * 4 + 64;
  if (cudaFuncSetAttribute((const void*)tcf_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return check_launch("conv_pf(cudaFuncSetAttribute)");
  const unsigned grid = (unsigned)(total_items < num_sms ? total_items : num_sms);
  if (getenv("ESM_DEBUG_PLAN"))
    fprintf(stderr, "[esm tcf] Cout=%d ncg=%d KD=%d K=%d phases=%d geom=(%d,%d,%d) mode=%d R=%d NPART=%d NT=%d x%d stages=%d (%zu B) items=%d grid=%u\n",
            d->Cout, k.ncg, k.KD, k.K, k.nphase, g0.Dp, g0.Hp, g0.P, k.mode, k.R, k.NPART, k.NT, k.ncot, ns, sb, k.total_items, grid);
  tcf_conv_kernel<<<grid, TF_THREADS, smem, (cudaStream_t)stream>>>(k);
  ++tcf_launches;
  return check_launch("conv_pf");
}
