// Streamed-weight tensor-core path of the convolution family: an implicit GEMM on tcgen05.mma (kind::tf32,
// accumulators in TMEM, fp32-grade split-TF32) with the kernel taps in K and the weights streamed through the
// operand ring by TMA bulk copies.  Replaces BasicConv (submodule.py:12-38) for the layers conv_tc.cu cannot
// hold resident weights for or has no gather for: the wide hourglass levels of `aggregation` (40->40, 72->72,
// ESMStereo.py:129-182), its stride-2 convs and its ConvTranspose3d k4 s2 p1 layers (as 8 sub-pixel phases of
// 2x2x2-tap convolutions), and the deep small-resolution 2D convs / deconvs of FeatUp (ESMStereo.py:79-125).
//
//   acc[m, co] = sum_{tap, ci} X[voxel(m) * stride + tap - pad, ci] * W[tap, ci, co]
//
// M = 128 consecutive voxels of the (flattened) output lattice, N = NT <= 128 output channels, one
// M128 x N x K8 MMA triple (hi*hi + lo*hi + hi*lo) per (8-channel group, tap).  With N this small an MMA is bound
// by the shared-memory read of A (46 clk, scratch/umma_test.cu), so the kernel wins where the FP32 pipe is starved
// instead: few voxels, many channels.
//
// Warp roles (704 threads, 1 CTA/SM, persistent): warps 0-3 epilogue (one per TMEM lane quadrant: TMEM -> BN /
// activation -> global), warp 4 MMA issuer (warp-uniform, one elected lane, TS form: A from tensor memory), warp 5
// weight streamer (one lane: one bulk copy per channel group, completing on the stage's `full` barrier), warps 6-21
// A-operand producers in two sets that take alternate stages (predicated global loads split into hi / lo and written
// to TENSOR MEMORY with tcgen05.st: row m of the tile at TMEM lane m).  A ring stage holds
// G = GC x GD x KW^2 (channel group, tap) slabs -- 9 for k3, 8 for the transposed layers -- with every tap offset
// a compile-time constant: the fence.proxy.async that publishes the producers' shared-memory stores compiles to
// MEMBAR.ALL.CTA, which also waits for the loads already in flight for the next stage, so a stage pays one global
// load latency (measured 844 clk per slab with one-slab stages), and a generic per-slab tap cursor unrolled nine
// times was 2900 instructions per stage (instruction-cache bound).
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace esm {

struct TcgK {
  esm_src_t src[3];
  int nsrc;
  int B, Cin, ncg;
  int Din, Hin, Win;            // input extent
  int Jd, Jh, Jw;               // output lattice per phase (transposed: ceil(out / 2))
  int Dout, Hout, Wout, Cout;
  int KD, KH, KW, taps;         // taps per phase
  int sz, sxy;                  // input step per lattice step (stride; 1 for transposed)
  int od, oh, ow;               // input offset of tap 0 at lattice 0 (regular: -pad; transposed: phase - 1, added per phase)
  int transposed, phases, phases_d;
  const float* wtc;             // split weight slabs (TcgPack)
  int CoutX;
  const float* scale;
  const float* shift;
  int act, act2;
  const float* out_mul;
  long long omB, omC, omH;
  const float* residual;
  float out_scale;
  float* out;
  long long oB, oC, oD, oH;
  int NT, ncot, mtiles, items, ctas, nstages, npass;
};

constexpr int TG_NEW = 4;                      // epilogue warps
constexpr int TG_MMA_WARP = TG_NEW;            // MMA issuer
constexpr int TG_W_WARP = TG_NEW + 1;          // weight streamer
constexpr int TG_PROD_WARP = TG_NEW + 2;       // first A producer warp
constexpr int TG_NTW = 8;                      // A producer warps per set: (lane quadrant, K half)
constexpr int TG_NSET = 2;                     // producer sets; set s fills the stages n with n % 2 == s
constexpr int TG_THREADS = 32 * (TG_NEW + 2 + TG_NSET * TG_NTW);

struct TgItem {
  int b, phase, mt, cot;
};
__device__ __forceinline__ TgItem tg_decode(const TcgK& p, int item) {
  TgItem t;
  t.cot = item % p.ncot;  // channel tiles of one voxel tile are neighbours: they share the A rows in L2
  int r = item / p.ncot;
  t.mt = r % p.mtiles;
  r /= p.mtiles;
  t.phase = r % p.phases;
  t.b = r / p.phases;
  return t;
}

// KW = KH: 3 (k3) or 2 (sub-pixel phase of a transposed k4 s2).  A stage covers GC channel groups x GD depth taps x
// KW x KW taps; slab g = ((gc * GD + gd) * KW + th) * KW + tw.
template <int KW, int GD, int GC>
__global__ void __launch_bounds__(TG_THREADS, 1) tcg_conv_kernel(const __grid_constant__ TcgK p) {
  constexpr int G = GC * GD * KW * KW;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef TC_PROFILE
  const long long tc_t0 = clock64();
#endif
  const int NT = p.NT, NS = p.nstages;
  const uint32_t BSLAB = (uint32_t)NT * 64;           // bytes of one B slab: [NT / 8][hi | lo][k / 4][8][4] floats
  const uint32_t STAGE = (uint32_t)G * BSLAB;         // shared memory holds only the B slabs of a stage
  // The A slabs live in TENSOR MEMORY (TS-form MMA): slab g of ring stage s = 16 columns (hi k0..7 | lo k0..7) at
  // TG_ACOL + (s * G + g) * 16.  In shared memory they cost 8 KB of stores + 12 KB of MMA reads per slab and the
  // kernel sat at the 128 B/clk shared-memory limit (30 KB per slab = 234 clk against 3 x 46 clk of MMA).
  constexpr uint32_t TG_ACOL = 512 - 2 * G * 16;      // two ring stages of A: the top 288 (k3) / 256 (transposed) columns
  constexpr int TG_ACC = (int)TG_ACOL;                // columns left for the accumulators (one buffer)
  uint8_t* s_stage = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * STAGE);
  uint64_t* empty = full + NS;
  uint64_t* accf = empty + NS;
  uint64_t* acce = accf + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acce + 2);
  const int ZST = p.KD / GD;                             // depth-tap stages per channel-group block
  const int SPI = ((p.ncg + GC - 1) / GC) * ZST;         // ring stages per item; the last block may hold fewer groups
  const int voxels = p.Jd * p.Jh * p.Jw;
  // Accumulator layout (one buffer of TG_ACC = 224 / 256 TMEM columns; the rest of TMEM is the A ring).  The tensor core truncates its fp32 accumulator on
  // every accumulate, a bias that grows with the chain length: with all 3 * ncg * taps MMAs of an item chained into
  // one accumulator the 240->240 layer was 7x less accurate than the fp32 pipe (feature error 1.4e-5 vs 2e-6).  So
  // the hi*hi products -- the only ones whose magnitude matters -- go round-robin by stage into P partial accumulators
  // (chains of ~27 MMAs, like the resident-weight engine's), the two correction products into one more, and the
  // epilogue adds the P + 1 columns of a channel in fp32 with round-to-nearest.
  const int P = max(1, min(12, TG_ACC / NT - 1));
  const int Pe = min(P, SPI);  // partials an item actually writes

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      tc_mbar_init(&full[i], TG_NTW + 1);  // the 8 producer warps of one set + the weight streamer's expect_tx arrive
      tc_mbar_init(&empty[i], 1);
    }
    tc_mbar_init(&accf[0], 1);
    tc_mbar_init(&acce[0], TG_NEW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp >= TG_PROD_WARP) {
    // ============================ A-operand producers ============================
    // Two sets of 8 warps take alternate ring stages.  A stage costs its producers one exposed global-load latency
    // (the MEMBAR inside fence.proxy.async waits for every load in flight, so a second register buffer cannot hide
    // it: measured 2.5k clk per 9-slab stage against 1.2k clk of MMAs); with two sets, one set's latency overlaps
    // the other set's conversion and stores.
    const int tw = (warp - TG_PROD_WARP) & (TG_NTW - 1);
    const int set = (warp - TG_PROD_WARP) / TG_NTW;
    const int q = warp & 3;        // TMEM lane quadrant this warp may access = rows q*32 .. q*32+31 of the tile
    const int khalf = tw >> 2;     // which 4 of the 8 channels of a group (each quadrant occurs twice among 8 consecutive warps)
    const int m = q * 32 + lane;   // A row
    // load cursor: (item, cg0, td0) names a stage
    int item = blockIdx.x, cg0 = 0, td0 = 0, cur_b = 0;
    int iz0 = 0, iy0 = 0, ix0 = 0;  // input coordinate of tap (0,0,0) for this thread's voxel
    bool vox_ok = false;
    auto enter_item = [&]() {
      if (item >= p.items) return;
      const TgItem ti = tg_decode(p, item);
      const int v = ti.mt * 128 + m;
      vox_ok = v < voxels;
      const int jx = v % p.Jw, r = v / p.Jw;
      const int jy = r % p.Jh, jz = r / p.Jh;
      const int pzw = ti.phase & 1, pzh = (ti.phase >> 1) & 1, pzd = (p.phases_d == 2) ? ((ti.phase >> 2) & 1) : 0;
      iz0 = jz * p.sz + p.od + (p.transposed ? pzd : 0);
      iy0 = jy * p.sxy + p.oh + (p.transposed ? pzh : 0);
      ix0 = jx * p.sxy + p.ow + (p.transposed ? pzw : 0);
      cur_b = ti.b;
      cg0 = td0 = 0;
    };
    // issues the loads of one stage (all of them before any use) and steps the cursor; returns the slabs loaded
    auto load_stage = [&](float (&v)[G][4]) -> int {
      if (item >= p.items) return 0;
      const int ngc = min(GC, p.ncg - cg0);
      bool xok[KW], yok[KW];
#pragma unroll
      for (int t = 0; t < KW; ++t) {
        xok[t] = (unsigned)(ix0 + t) < (unsigned)p.Win;
        yok[t] = (unsigned)(iy0 + t) < (unsigned)p.Hin;
      }
#pragma unroll
      for (int gc = 0; gc < GC; ++gc) {
        if (gc < ngc) {
          int rel = (cg0 + gc) * 8 + khalf * 4, k = 0;
          while (k < p.nsrc - 1 && rel >= p.src[k].C) {  // host guarantees 8-channel groups never straddle sources
            rel -= p.src[k].C;
            ++k;
          }
          const esm_src_t& sr = p.src[k];
          const int sC = (int)sr.sC, sD = (int)sr.sD, sH = (int)sr.sH;
          const int nch = sr.C - rel;  // valid channels from `rel` on (<= 0 past the last group)
          const float* bp = sr.ptr + (long long)cur_b * sr.sB;
          const int off0 = rel * sC + (iz0 + td0) * sD + iy0 * sH + ix0;
#pragma unroll
          for (int gd = 0; gd < GD; ++gd) {
            const bool zok = vox_ok && (unsigned)(iz0 + td0 + gd) < (unsigned)p.Din;
#pragma unroll
            for (int th = 0; th < KW; ++th) {
#pragma unroll
              for (int tw = 0; tw < KW; ++tw) {
                const bool ok = zok && yok[th] && xok[tw];
                const int off = off0 + gd * sD + th * sH + tw;
#pragma unroll
                for (int c = 0; c < 4; ++c) v[((gc * GD + gd) * KW + th) * KW + tw][c] = (ok && c < nch) ? __ldg(bp + (off + c * sC)) : 0.f;
              }
            }
          }
        }
      }
      return ngc * (GD * KW * KW);
    };
    auto next_stage = [&]() {
      if (item >= p.items) return;
      td0 += GD;
      if (td0 >= p.KD) {
        td0 = 0;
        cg0 += GC;
        if (cg0 >= p.ncg) {
          item += p.ctas;
          enter_item();
        }
      }
    };
    uint32_t st = (uint32_t)set % (uint32_t)NS, ph = (uint32_t)set / (uint32_t)NS;  // ring slot / phase of stage n = set
    auto store_stage = [&](const float (&v)[G][4], int nv) {
      tc_mbar_wait(&empty[st], ph ^ 1, 200 + (int)st);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + TG_ACOL + st * (G * 16) + khalf * 4;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (g < nv) {
          const float h0 = tc_rna(v[g][0]), h1 = tc_rna(v[g][1]), h2 = tc_rna(v[g][2]), h3 = tc_rna(v[g][3]);
          tc_st4(ta + g * 16, h0, h1, h2, h3);
          if (p.npass == 3) tc_st4(ta + g * 16 + 8, tc_lo(v[g][0], h0), tc_lo(v[g][1], h1), tc_lo(v[g][2], h2), tc_lo(v[g][3], h3));
        }
      }
      tc_st_wait();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&full[st]);
      st += TG_NSET;  // this set's next stage is TG_NSET further along the ring
      while (st >= (uint32_t)NS) {
        st -= (uint32_t)NS;
        ph ^= 1;
      }
    };
    float v[G][4];
    enter_item();
    for (int i = 0; i < set; ++i) next_stage();
    for (;;) {
      const int nv = load_stage(v);
      if (nv == 0) break;
#pragma unroll
      for (int i = 0; i < TG_NSET; ++i) next_stage();
      store_stage(v, nv);
    }
  } else if (warp == TG_W_WARP) {
    // ============================ weight streamer ============================
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      const long long slab_floats = 16ll * p.CoutX;
      constexpr int TPB = GD * KW * KW;  // taps of one channel group in a stage: consecutive slabs of the pack
      for (int item = blockIdx.x; item < p.items; item += p.ctas) {
        const TgItem ti = tg_decode(p, item);
        const int n0 = ti.cot * NT;
        const int rows = min(NT, p.CoutX - n0);
        const uint32_t bytes = (uint32_t)rows * 64;
        const bool whole = p.ncot == 1;  // the tile is the whole slab: the TPB slabs of a group are one contiguous copy
        for (int cg0 = 0; cg0 < p.ncg; cg0 += GC) {
          const int ngc = min(GC, p.ncg - cg0);
          for (int td0 = 0; td0 < p.KD; td0 += GD) {
            tc_mbar_wait(&empty[st], ph ^ 1, 300 + (int)st);
            tc_mbar_expect_tx(&full[st], (uint32_t)(ngc * TPB) * bytes);
            uint8_t* sb = s_stage + (size_t)st * STAGE;
            for (int gc = 0; gc < ngc; ++gc) {
              const float* src = p.wtc + (((long long)ti.phase * p.ncg + cg0 + gc) * p.taps + td0 * KW * KW) * slab_floats + n0 * 16;
              if (whole) {
                tc_bulk_g2s(sb + (size_t)(gc * TPB) * BSLAB, src, TPB * bytes, &full[st]);
              } else {
                for (int t = 0; t < TPB; ++t) tc_bulk_g2s(sb + (size_t)(gc * TPB + t) * BSLAB, src + t * slab_floats, bytes, &full[st]);
              }
            }
            if (++st == (uint32_t)NS) {
              st = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == TG_MMA_WARP) {
    // ============================ MMA issuer (warp-uniform, see conv_tc.cu) ============================
    const uint32_t leader = tc_elect();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    // D = f32, A = B = tf32, both K-major, N = NT, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t b0 = tc_desc(tc_smem_u32(s_stage), 128, 512);
    const bool three = p.npass == 3;
    uint32_t st = 0, ph = 0, ai = 0;
    for (int item = blockIdx.x; item < p.items; item += p.ctas) {
      const uint32_t ab = 0, aph = ai & 1;
      tc_mbar_wait(&acce[ab], aph ^ 1, 400 + (int)ab);  // epilogue has drained the accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t dbuf = tmem_u, dcorr = dbuf + P * NT;
      int part = 0;
      for (int sg = 0; sg < SPI; ++sg) {
        const int nv = min(GC, p.ncg - (sg / ZST) * GC) * (GD * KW * KW);
        tc_mbar_wait(&full[st], ph, 500 + (int)st);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) {
          const uint64_t so = (uint64_t)((st * STAGE) >> 4);
          for (int g = 0; g < nv; ++g) {
            const uint32_t a_hi = tmem_u + TG_ACOL + (st * G + g) * 16;
            const uint64_t b_hi = b0 + so + (uint64_t)((g * BSLAB) >> 4);
            tc_mma_ts(dbuf + part * NT, a_hi, b_hi, idesc, (sg >= P || g) ? 1u : 0u);
            if (three) {
              tc_mma_ts(dcorr, a_hi + 8, b_hi, idesc, (sg | g) ? 1u : 0u);
              tc_mma_ts(dcorr, a_hi, b_hi + (256 >> 4), idesc, 1u);
            }
          }
          tc_commit(&empty[st]);
        }
        __syncwarp();
        if (++part == P) part = 0;
        if (++st == (uint32_t)NS) {
          st = 0;
          ph ^= 1;
        }
      }
      if (leader) tc_commit(&accf[ab]);
      __syncwarp();
      ++ai;
    }
  } else {
    // ============================ epilogue ============================
    const int q = warp;  // TMEM lane quadrant
    const int m = q * 32 + lane;
    const bool post = p.out_mul || p.residual || p.act2 != ESM_ACT_NONE;
    const bool gelu = p.act == ESM_ACT_GELU;
    const int so = p.transposed ? 2 : 1;
    uint32_t ai = 0;
    for (int item = blockIdx.x; item < p.items; item += p.ctas) {
      const TgItem ti = tg_decode(p, item);
      const int v = ti.mt * 128 + m;
      const int jx = v % p.Jw, r = v / p.Jw;
      const int jy = r % p.Jh, jz = r / p.Jh;
      const int pzw = ti.phase & 1, pzh = (ti.phase >> 1) & 1, pzd = (p.phases_d == 2) ? ((ti.phase >> 2) & 1) : 0;
      const int oz = p.transposed ? jz * (p.phases_d == 2 ? 2 : 1) + pzd : jz;
      const int oy = p.transposed ? jy * so + pzh : jy, ox = p.transposed ? jx * so + pzw : jx;
      const bool ok = v < voxels && oz < p.Dout && oy < p.Hout && ox < p.Wout;
      const int co0 = ti.cot * NT;
      const long long obase = (long long)ti.b * p.oB + (long long)oz * p.oD + (long long)oy * p.oH + ox;
      float* op = p.out + obase;
      const float* rp = p.residual ? p.residual + obase : nullptr;
      const float* mp = p.out_mul ? p.out_mul + (long long)ti.b * p.omB + (long long)oy * p.omH + ox : nullptr;
      const uint32_t ab = 0, aph = ai & 1;
      tc_mbar_wait(&accf[ab], aph, 600 + (int)ab);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16);
      const int nacc = Pe + (p.npass == 3 ? 1 : 0);  // columns to add per channel: partials 0..Pe-1, then the corrections at P
      // expected-value correction of the accumulator truncation: each of the ~slabs/Pe chained accumulates into a
      // partial rounds it toward zero, -1.25e-8 relative per step on average (scripts/engine_accuracy.py: the bias
      // column; the FP32 pipe rounds to nearest and has none)
      const float debias = 1.0f + TC_TRUNC_BIAS * (float)(p.ncg * p.taps) / (float)Pe;
      for (int c8 = 0; c8 < NT; c8 += 8) {
        float rv[8], t0[8], t1[8];
        tc_ld8(tb + c8, rv);
        for (int j = 1; j < nacc; j += 2) {
          const bool two = j + 1 < nacc;
          tc_ld8(tb + (j < Pe ? j : P) * NT + c8, t0);
          if (two) tc_ld8(tb + (j + 1 < Pe ? j + 1 : P) * NT + c8, t1);
          tc_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) rv[i] += two ? t0[i] + t1[i] : t0[i];
        }
        tc_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] *= debias;
        if (c8 + 8 >= NT) {  // last TMEM read of this item: hand the accumulator buffer back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) tc_mbar_arrive(&acce[ab]);
        }
        const int co = co0 + c8;
        if (ok && co < p.Cout) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int cj = min(co + j, p.Cout - 1);
            rv[j] = fmaf(rv[j], p.scale ? __ldg(p.scale + cj) : 1.f, p.shift ? __ldg(p.shift + cj) : 0.f);
          }
          if (gelu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = tc_gelu(rv[j]);
          } else if (p.act != ESM_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
              const float4 a = apply_act4(make_float4(rv[j], rv[j + 1], rv[j + 2], rv[j + 3]), p.act);
              rv[j] = a.x; rv[j + 1] = a.y; rv[j + 2] = a.z; rv[j + 3] = a.w;
            }
          }
          if (post) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (co + j < p.Cout) {
                if (mp) rv[j] *= __ldg(mp + (long long)(co + j) * p.omC);
                if (rp) rv[j] += __ldg(rp + (long long)(co + j) * p.oC);
              }
            }
            if (p.act2 != ESM_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 8; j += 4) {
                const float4 a = apply_act4(make_float4(rv[j], rv[j + 1], rv[j + 2], rv[j + 3]), p.act2);
                rv[j] = a.x; rv[j + 1] = a.y; rv[j + 2] = a.z; rv[j + 3] = a.w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (co + j < p.Cout) op[(long long)(co + j) * p.oC] = rv[j] * p.out_scale;
        }
      }
      ++ai;
    }
  }
#ifdef TC_PROFILE
  if (blockIdx.x == 0 && lane == 0) {
    printf("tcg_prof warp %2d: done at %8lld clk; waits: empty %8llu (w %8llu)  acce %8llu  full %8llu  accf %8llu\n", warp, clock64() - tc_t0,
           tc_prof_wait[warp][2], tc_prof_wait[warp][3], tc_prof_wait[warp][4], tc_prof_wait[warp][5], tc_prof_wait[warp][6]);
    for (int i = 0; i < 8; ++i) tc_prof_wait[warp][i] = 0;
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static long long tcg_launches = 0;

bool tcg_conv_plan(const esm_conv_t* d, int num_sms, int npass, TcgPlan* plan) {
  if (d->src_mode != ESM_SRC_TENSORS || d->pixel_shuffle || d->in_mul) return false;
  if (d->Cin < 8 || num_sms <= 0) return false;
  if (d->stride != 1 && d->stride != 2) return false;
  {  // ESM_TCG_DIMS=2 / 3: only 2D / only 3D layers (accuracy experiments, scripts/sceneflow_parity.py)
    const char* dims = getenv("ESM_TCG_DIMS");
    const bool is3d = d->Dout > 1 || d->Din > 1;
    if (dims && ((dims[0] == '2' && is3d) || (dims[0] == '3' && !is3d))) return false;
  }
  if (d->transposed && !(d->kh == 4 && d->kw == 4 && (d->kd == 4 || d->kd == 1) && d->stride == 2)) return false;
  if (!d->transposed && !(d->kh == 3 && d->kw == 3 && (d->kd == 3 || d->kd == 1))) return false;  // k1 / k5 stay on the other engines
  for (int i = 0; i + 1 < d->nsrc; ++i)
    if (d->src[i].C % 8) return false;
  // the producers address one batch item of one source with 32-bit element offsets
  for (int i = 0; i < d->nsrc; ++i)
    if ((long long)d->src[i].C * d->src[i].sC >= (1ll << 31) || (long long)d->Din * d->src[i].sD >= (1ll << 31)) return false;
  const TcgPack tp = tcg_pack_geom(d->Cout, d->Cin, d->kd, d->kh, d->kw, d->transposed);
  if (tp.elems == 0) return false;
  const int phases_d = (d->transposed && d->kd == 4) ? 2 : 1;
  const int Jd = d->transposed ? ceil_div(d->Dout, phases_d) : d->Dout;
  const int Jh = d->transposed ? ceil_div(d->Hout, 2) : d->Hout;
  const int Jw = d->transposed ? ceil_div(d->Wout, 2) : d->Wout;
  const long long voxels = (long long)Jd * Jh * Jw;
  if (voxels >= (1ll << 30)) return false;
  const int mtiles = (int)ceil_div_ll(voxels, 128);
  const long long mt_all = (long long)d->B * tp.phases * mtiles;
  const int KT = tp.ncg * tp.taps;
  // channel tile: the MMA costs ~max(46, N/2 + 10) clk (A-read bound below N ~ 72), a CTA walks KT slabs per item;
  // pick the tile count that finishes first
  // Accuracy constraint: the tensor core truncates its accumulator at every accumulate, a systematic toward-zero bias
  // of ~1.2e-8 (relative) per chained hi*hi MMA (scripts/engine_accuracy.py).  The chain of a partial accumulator is
  // slabs / P stages' worth of MMAs with P = columns / NT - 1, so narrower channel tiles buy shorter chains: tiles whose
  // chain exceeds TG_MAXCHAIN slabs are skipped (FeatUp's 240 -> 240 at NT = 64 chained 135 and, upstream of the cost
  // volume, doubled the top-2 flips of the SceneFlow-shape test).
  constexpr int TG_MAXCHAIN = 160;
  const int acc_cols = 512 - 2 * (d->transposed ? 8 : 9) * 16;
  double best = 1e30;
  int best_ncot = 0, best_nt = 0;
  for (int ncot = 1; ncot <= 32; ++ncot) {
    const int nt = round_up(ceil_div(tp.CoutX, ncot), 8);
    if (nt > 112 || (ncot > 1 && (ncot - 1) * nt >= tp.CoutX)) continue;  // one partial + the corrections in 224 columns
    const int P = acc_cols / nt - 1 > 12 ? 12 : acc_cols / nt - 1;
    if (ceil_div(KT, P) > TG_MAXCHAIN && nt > 8) continue;
    const long long items = mt_all * ncot;
    const long long waves = ceil_div_ll(items, num_sms);
    const double mma = (nt / 2.0 + 10.0) > 46.0 ? (nt / 2.0 + 10.0) : 46.0;
    const double cost = (double)waves * (KT * 3.0 * mma + 600.0 + nt * 12.0);
    if (cost < best) {
      best = cost;
      best_ncot = ncot;
      best_nt = nt;
    }
  }
  if (!best_ncot) return false;
  plan->NT = best_nt;
  plan->ncot = best_ncot;
  plan->mtiles = mtiles;
  plan->npass = npass;
  const int G = d->transposed ? 8 : 9;
  plan->nstages = 2;  // the A ring in tensor memory has two stages; the B ring in shared memory follows it
  plan->smem = (size_t)plan->nstages * G * best_nt * 64 + 1024;
  const long long items = mt_all * best_ncot;
  if (items >= (1ll << 31)) return false;
  plan->ctas = (int)(items < num_sms ? items : num_sms);
  return true;
}

int tcg_conv_launch(const esm_conv_t* d, const TcgPlan& plan, cudaStream_t st) {
  TcgK k;
  memset(&k, 0, sizeof(k));
  for (int i = 0; i < d->nsrc; ++i) k.src[i] = d->src[i];
  k.nsrc = d->nsrc;
  const TcgPack tp = tcg_pack_geom(d->Cout, d->Cin, d->kd, d->kh, d->kw, d->transposed);
  k.B = d->B;
  k.Cin = d->Cin;
  k.ncg = tp.ncg;
  k.Din = d->Din;
  k.Hin = d->Hin;
  k.Win = d->Win;
  k.phases_d = (d->transposed && d->kd == 4) ? 2 : 1;
  k.Jd = d->transposed ? ceil_div(d->Dout, k.phases_d) : d->Dout;
  k.Jh = d->transposed ? ceil_div(d->Hout, 2) : d->Hout;
  k.Jw = d->transposed ? ceil_div(d->Wout, 2) : d->Wout;
  k.Dout = d->Dout;
  k.Hout = d->Hout;
  k.Wout = d->Wout;
  k.Cout = d->Cout;
  k.KD = tp.KD;
  k.KH = tp.KH;
  k.KW = tp.KW;
  k.taps = tp.taps;
  k.transposed = d->transposed;
  k.phases = tp.phases;
  if (d->transposed) {
    k.sz = k.sxy = 1;
    k.od = (k.phases_d == 2) ? -1 : 0;  // input index of tap t at lattice j, phase pz: j + pz - 1 + t
    k.oh = k.ow = -1;
  } else {
    k.sz = d->kd == 1 ? 1 : d->stride;
    k.sxy = d->stride;
    k.od = -d->pd;
    k.oh = -d->ph;
    k.ow = -d->pw;
  }
  k.wtc = d->weight + tp.offset;
  k.CoutX = tp.CoutX;
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
  k.out = d->out;
  k.oB = d->oB;
  k.oC = d->oC;
  k.oD = d->oD;
  k.oH = d->oH;
  k.NT = plan.NT;
  k.ncot = plan.ncot;
  k.mtiles = plan.mtiles;
  k.items = d->B * tp.phases * plan.mtiles * plan.ncot;
  k.ctas = plan.ctas;
  k.nstages = plan.nstages;
  k.npass = plan.npass;
  void (*fn)(const TcgK) = !d->transposed ? tcg_conv_kernel<3, 1, 1> : (d->kd == 4 ? tcg_conv_kernel<2, 2, 1> : tcg_conv_kernel<2, 1, 2>);
  if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return check_launch("conv(tcg, cudaFuncSetAttribute)");
  fn<<<(unsigned)plan.ctas, TG_THREADS, plan.smem, st>>>(k);
  ++tcg_launches;
  return check_launch("conv(tcg)");
}

}  // namespace esm

extern "C" long long esm_tcg_conv_launches(void) { return esm::tcg_launches; }
