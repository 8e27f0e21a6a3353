// Tensor-core path of the convolution family: k3 s1 p1 convolutions (2D and 3D) as implicit GEMMs on
// tcgen05.mma (kind::tf32, accumulators in TMEM), with fp32-grade accuracy from a split-TF32 scheme.
// Replaces BasicConv (submodule.py:12-38) for group_stem / agg / the k3 layers of `aggregation`
// (ESMStereo.py:129-182, 620-622) and the k3 s1 2D convs; with GWC also build_gwc_volume
// (submodule.py:151-161) fused in front of group_stem.
//
// Shape problem and the answer to it.  Every GEMM here has a tiny Cout (8..24 per CTA) and the voxel
// dimension must be M, so an SS-mode MMA is bound by the shared-memory read of the A operand: measured
// 46 clk per M128 x K8 dispatch whatever N <= 64 is (scratch/umma_test.cu).  Re-reading the im2col rows
// once per tap at N = Cout would make the tensor pipe slower than the FP32 pipe.  So the taps go into N
// instead of K:
//     acc[z_o][y_in][x_in, (kh,kw,co)] = sum_{kd,ci} X[z_o+kd-1, y_in, x_in, ci] * W[kd,kh,kw,ci,co]
// i.e. one M128 x N(9*COT) x K8 MMA per (input row, kd, 8-channel group) -- N = 72..216, at or near the
// MMA's own rate -- and the 9-way (kh,kw) gather that remains,
//     out[z_o, y, x, co] = sum_{kh,kw} acc[z_o][y+kh-1][x+kw-1, (kh,kw,co)],
// is done by the epilogue warps on CUDA cores: kw by two warp shuffles, kh by a rolling 3-row register
// window while the CTA marches down y.  A CTA tile is 4 "strips" (one per epilogue warp = TMEM lane
// quadrant): 32 consecutive input columns (30 outputs + halo) of one image row each.
//
// Accuracy.  kind::tf32 truncates the operands to 10 mantissa bits (measured).  Operands are split
// x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi) and three MMAs accumulate hi*hi + lo*hi +
// hi*lo in fp32 (relative error ~2^-22 per product, measured 5e-7 worst case against fp64 on K=8 dot
// products).  npass = 1 keeps only hi*hi (single-pass TF32 fast mode).
//
// Warp roles (416 threads, 1 CTA/SM, persistent): warps 0-3 epilogue (TMEM -> registers -> global),
// warps 8-10 MMA issuers (one thread each, one per output plane), warps 11-18 operand producers: they load fp32 activations (or build
// the correlation from left/right feature rows), split them and write the K-major, non-swizzled UMMA
// operand tiles.  Three mbarrier pipelines: operand ring full/empty, accumulator full/empty.
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace esm {

struct TcK {
  esm_src_t src[3];
  int nsrc, cpg;
  int B, Cin, ncg, D, H, W;
  int Cout, CinPad, CoutPad;
  const float* weight;  // fp32 pack of esm_pack_conv_weight_f32: [tap][CinPad][CoutPad]
  const float* wimg;    // tck_conv_kernel: the split weight image of the pack (TcImg)
  int tail_help;        // tck_conv_kernel: the producer warps help with the epilogue of the CTA's last rows
  const float* scale;
  const float* shift;
  int act, act2;
  const float* out_mul;
  long long omB, omC, omH;
  const float* residual;
  float out_scale;
  float* out;
  long long oB, oC, oD, oH;
  int nseg, segw, rows, nstrips, groups, ztiles;
  int items_per_cot, ctas_per_cot, nstages, npass;
  int ps;  // PixelShuffle(2) store (pointwise layers only): channel co -> out[co / 4][2y + (co / 2) % 2][2x + co % 2]
};

constexpr int TC_NTW = 8;                               // operand-producer warps
// Epilogue warps NEW (template): 8 = two per TMEM lane quadrant, half the channel tile each -- 608 threads with the
// three MMA warps (one per output plane z_o, TZ of them active; ptxas sizes for 640: 96 registers); 16 = four per
// quadrant, a quarter of the tile each, one MMA warp (TZ = 1 only): 800 threads at 80 registers.
__host__ __device__ constexpr int tc_nmw(int NEW) { return NEW == 8 ? 3 : 1; }
__host__ __device__ constexpr int tc_threads(int NEW) { return 32 * (NEW + tc_nmw(NEW) + TC_NTW); }


struct TcItem {
  int b, z0, grp;
};
__device__ __forceinline__ TcItem tc_decode(const TcK& p, int item, int TZ) {
  TcItem t;
  t.grp = item % p.groups;
  const int r = item / p.groups;
  t.z0 = (r % p.ztiles) * TZ;
  t.b = r / p.ztiles;
  return t;
}

// TAPS = 9: k3 s1 p1 in (h, w) as described above.  TAPS = 1: pointwise (k1) convolution -- the same
// pipeline without halos, shuffles or the rolling window (N = COT, 32 output columns per strip).
template <int COT, int TZ, int KD, bool GWC, int TAPS = 9, int NEW = 8>
__global__ void __launch_bounds__(tc_threads(NEW), 1) tc_conv_kernel(const __grid_constant__ TcK p) {
  constexpr int TC_NEW = NEW, TC_NMW = tc_nmw(NEW), TC_MMA_WARP = TC_NEW, TC_PROD_WARP = TC_NEW + TC_NMW;
  static_assert(TZ <= TC_NMW, "one MMA warp per output plane");
  static_assert(COT % (NEW / 4 * 4) == 0, "whole 4-channel units per epilogue warp");
  constexpr int NB = TAPS * COT;               // accumulator columns per (z_o, y_in) row tile
  constexpr int HALO = TAPS == 9 ? 1 : 0;
  constexpr int NROW = TZ + KD - 1;            // input planes per stage
  constexpr int CGS = (NROW == 1) ? 4 : 1;     // 8-channel groups per stage
  constexpr int TPW = CGS * NROW;              // row tiles per stage = tasks per producer warp
  constexpr int ROW_BYTES = 8192;              // row tile: hi [2][128][4] floats, then lo [2][128][4]
  constexpr int STAGE_BYTES = TPW * ROW_BYTES;
  constexpr int WSLAB = NB * 32;               // one (cg, kd, hi|lo) B operand: [2][NB][4] floats
  constexpr int ACC_COLS = 256;                // TMEM columns per accumulator buffer (TZ*NB <= 256)
  constexpr int CW = COT / (NEW / 4);          // output channels per epilogue warp
  static_assert(TZ * NB <= ACC_COLS, "accumulator does not fit");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef TC_PROFILE
  const long long tc_t0 = clock64();
#endif
  const int ncg = p.ncg;
  const int NS = p.nstages;
  const uint32_t wbytes = (uint32_t)ncg * KD * 2 * WSLAB;
  uint8_t* s_w = smem;
  uint8_t* s_stage = smem + ((wbytes + 127u) & ~127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_stage + (size_t)NS * STAGE_BYTES);
  uint64_t* empty = full + NS;
  uint64_t* accf = empty + NS;
  uint64_t* acce = accf + 2;
  uint64_t* wready = acce + 2;  // the resident weights are staged (one arrival per staging warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wready + 1);
  float* s_aff = reinterpret_cast<float*>(tmem_slot + 2);  // [2][COT] scale, shift of this channel tile
  const int cot = blockIdx.x / p.ctas_per_cot;
  const int cta = blockIdx.x % p.ctas_per_cot;
  const int nsteps = p.rows + 2 * HALO;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      tc_mbar_init(&full[i], TC_NTW);
      tc_mbar_init(&empty[i], TZ);  // one tcgen05.commit per MMA warp
    }
    for (int i = 0; i < 2; ++i) {
      tc_mbar_init(&accf[i], TZ);
      tc_mbar_init(&acce[i], TC_NEW);
    }
    tc_mbar_init(wready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 2 * COT) {
    const int c = tid % COT, co = cot * COT + c;
    const float* src = tid < COT ? p.scale : p.shift;
    s_aff[tid] = (src && co < p.Cout) ? __ldg(src + co) : (tid < COT ? 1.f : 0.f);
  }
  // Resident weights.  With an image in the pack (conv_tc.cuh TcImg) the channel tile is one bulk copy per (8-channel
  // group, kd) slab pair, issued below and awaited by the MMA issuers on `wready` while the producers already load.
  // Without one (the gwc stem: the group mean's 0.5 is folded into its weights) every thread stages its share here.
  if (!p.wimg) {
    // resident weights of this channel tile: split and laid out as UMMA B operands
    // (row n = co*TAPS + kh*3+kw: the 9 taps of a channel are adjacent accumulator columns; K = 8 channels of group cg)
    {
      const int total = ncg * KD * NB * 8;
      constexpr int U = 4;
      for (int base = tid; base < total; base += U * (int)blockDim.x) {
        float w[U];
        uint32_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = base + u * (int)blockDim.x;
          const int col = idx % COT;
          int t = idx / COT;
          const int k = t & 7;
          t >>= 3;
          const int tap2 = t % TAPS;
          t /= TAPS;
          const int kd = t % KD;
          const int cg = t / KD;
          const int co = cot * COT + col, ci = cg * 8 + k;
          w[u] = 0.f;
          if (idx < total && co < p.CoutPad && ci < p.CinPad)
            w[u] = (GWC ? 0.5f : 1.0f) * __ldg(p.weight + ((long long)(kd * TAPS + tap2) * p.CinPad + ci) * p.CoutPad + co);
          off[u] = (uint32_t)((cg * KD + kd) * 2) * WSLAB + (uint32_t)(k >> 2) * (NB * 16) + (uint32_t)(col * TAPS + tap2) * 16 + (k & 3) * 4;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (base + u * (int)blockDim.x < total) {
            const float hi = tc_rna(w[u]);
            *reinterpret_cast<float*>(s_w + off[u]) = hi;
            *reinterpret_cast<float*>(s_w + off[u] + WSLAB) = tc_lo(w[u], hi);
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (p.wimg) {
    if (tid == TC_MMA_WARP * 32) {
      tc_mbar_expect_tx(wready, wbytes);
      const float* img = p.wimg + (long long)cot * (wbytes / 4);
      for (int sl = 0; sl < ncg * KD; ++sl) tc_bulk_g2s(s_w + (size_t)sl * (2 * WSLAB), img + (long long)sl * (2 * WSLAB / 4), 2 * WSLAB, wready);
    }
  } else if (tid == TC_MMA_WARP * 32) {
    tc_mbar_arrive(wready);  // staged before the barrier above
  }
#ifdef TC_PROFILE
  const long long tc_t1 = clock64();
#endif

  if (warp >= TC_PROD_WARP) {
    // ============================ operand producers ============================
    const int tw = warp - TC_PROD_WARP;
    const int q = (tw >> 1) & 3;   // strip (= TMEM lane quadrant) this warp feeds
    const int khalf = tw & 1;      // which 4 of the 8 channels of a group
    const int m = q * 32 + lane;   // A row
    uint32_t st = 0, ph = 0;       // ring cursor: stage and its phase bit (no runtime division in the loop)
    // load cursor (item, step, cgb) and what it caches per item; offsets are 32-bit (host checks the extents)
    int item = cta, step = 0, cgb = 0;
    const float* base[3] = {nullptr, nullptr, nullptr};
    int z0 = 0, ya = 0, x = 0;
    bool strip_ok = false;
    auto enter_item = [&]() {
      if (item >= p.items_per_cot) return;
      const TcItem ti = tc_decode(p, item, TZ);
      const int strip = ti.grp * 4 + q;
      const int seg = strip % p.nseg, ys = strip / p.nseg;
      x = seg * p.segw + lane - HALO;
      ya = ys * p.rows;
      z0 = ti.z0;
      strip_ok = strip < p.nstrips && x >= 0 && x < p.W && lane < p.segw + 2 * HALO;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (i < p.nsrc) base[i] = p.src[i].ptr + (long long)ti.b * p.src[i].sB;
    };
    auto load = [&](float (&v)[TPW][4]) {
      const int y = ya - HALO + step;
      const bool ok = strip_ok && (unsigned)y < (unsigned)p.H;
      if (GWC) {
        // v = L[2g]*R[2g](x-d) + L[2g+1]*R[2g+1](x-d), un-contracted like the reference; the 0.5 of the group mean (submodule.py:147, cpg == 2) is
        // folded into the resident weights, exactly (a power of two).  All loads are issued first (predicated,
        // never branched around) so that they overlap.
        const int g0 = cgb * 8 + khalf * 4;
        const int nch = 2 * min(4, p.Cin - g0);  // valid feature channels of this half group
        const int sC = (int)p.src[0].sC;
        const int off = (g0 * 2) * sC + y * (int)p.src[0].sH + x;
        float l[8], rr[NROW][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) l[c] = (ok && c < nch) ? __ldg(base[0] + (off + c * sC)) : 0.f;
#pragma unroll
        for (int r = 0; r < NROW; ++r) {
          const int d = z0 + r - KD / 2;
          const bool okd = ok && (unsigned)d < (unsigned)p.D && x >= d;
#pragma unroll
          for (int c = 0; c < 8; ++c) rr[r][c] = (okd && c < nch) ? __ldg(base[1] + (off + c * sC - d)) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < NROW; ++r)
#pragma unroll
          for (int g = 0; g < 4; ++g) v[r][g] = __fadd_rn(__fmul_rn(l[2 * g], rr[r][2 * g]), __fmul_rn(l[2 * g + 1], rr[r][2 * g + 1]));
      } else {
#pragma unroll
        for (int cgl = 0; cgl < CGS; ++cgl) {
          int rel = (cgb + cgl) * 8 + khalf * 4, k = 0;
          if (p.nsrc > 1) {
            while (k < p.nsrc - 1 && rel >= p.src[k].C) {  // host guarantees 8-channel groups never straddle sources
              rel -= p.src[k].C;
              ++k;
            }
          }
          const int sC = (int)p.src[k].sC, sD = (int)p.src[k].sD;
          const int nch = p.src[k].C - rel;  // valid channels from `rel` on (<= 0 past the last group)
          const float* bp = p.nsrc > 1 ? (k == 0 ? base[0] : k == 1 ? base[1] : base[2]) : base[0];
          const int off = rel * sC + (z0 - KD / 2) * sD + y * (int)p.src[k].sH + x;
#pragma unroll
          for (int r = 0; r < NROW; ++r) {
            const bool okz = ok && (unsigned)(z0 + r - KD / 2) < (unsigned)p.D;
#pragma unroll
            for (int c = 0; c < 4; ++c) v[cgl * NROW + r][c] = (okz && c < nch) ? __ldg(bp + (off + r * sD + c * sC)) : 0.f;
          }
        }
      }
    };
    auto advance = [&]() {
      cgb += CGS;
      if (cgb >= ncg) {
        cgb = 0;
        if (++step >= nsteps) {
          step = 0;
          item += p.ctas_per_cot;
          enter_item();
        }
      }
    };
    auto store_stage = [&](const float (&v)[TPW][4]) {
      tc_mbar_wait(&empty[st], ph ^ 1, 200 + (int)st);
      uint8_t* sb = s_stage + (size_t)st * STAGE_BYTES + khalf * 2048 + m * 16;
#pragma unroll
      for (int j = 0; j < TPW; ++j) {
        float4 hi, lo;
        hi.x = tc_rna(v[j][0]);
        hi.y = tc_rna(v[j][1]);
        hi.z = tc_rna(v[j][2]);
        hi.w = tc_rna(v[j][3]);
        *reinterpret_cast<float4*>(sb + j * ROW_BYTES) = hi;
        if (p.npass == 3) {
          lo.x = tc_lo(v[j][0], hi.x);
          lo.y = tc_lo(v[j][1], hi.y);
          lo.z = tc_lo(v[j][2], hi.z);
          lo.w = tc_lo(v[j][3], hi.w);
          *reinterpret_cast<float4*>(sb + j * ROW_BYTES + 4096) = lo;
        }
      }
      // The proxy fence that makes these generic-proxy stores visible to the MMA sits on the CONSUMER side of the
      // release (arrive below) / acquire (the issuer's wait on `full`) chain: here it compiles to MEMBAR.ALL.CTA +
      // FENCE.VIEW.ASYNC, and the MEMBAR waited for the loads of the next stage already in flight -- one exposed
      // load latency per stage.  -DTC_FENCE_WRITER restores the writer-side fence (A/B: 1.970 -> 1.924 ms per KITTI pair
      // on one box, 1.965 -> 1.952 on another; tests/test_gpu_ops.py::test_tensor_core_resident_engine_is_race_free).
#ifdef TC_FENCE_WRITER
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA
#endif
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&full[st]);
      if (++st == (uint32_t)NS) {
        st = 0;
        ph ^= 1;
      }
    };
    // two register buffers, loop unrolled by two: the loads of stage i+1 are in flight while stage i is
    // converted and stored (no register copies, so nothing waits on a load before its own store)
    float va[TPW][4], vb[TPW][4];
    enter_item();
    if (item < p.items_per_cot) load(va);
    while (item < p.items_per_cot) {
      advance();
      if (item < p.items_per_cot) load(vb);
      store_stage(va);
      if (item >= p.items_per_cot) break;
      advance();
      if (item < p.items_per_cot) load(va);
      store_stage(vb);
    }
  } else if (warp >= TC_MMA_WARP) {
    // ============================ MMA issuers ============================
    // One thread per output plane z_o: the MMAs of different planes accumulate into different TMEM columns, so
    // their issue streams are independent -- and the issue stream of a single thread, not the tensor pipe, is
    // what bounded the kernel (measured: ~140 clk per MMA issued against 50 clk per MMA executed).  Every
    // issuer commits to the stage's `empty` barrier and to the step's `accf` barrier (count TZ).
    // The whole warp runs the loop so that every descriptor lives in uniform registers and tcgen05.mma issues
    // straight from them; with a single active lane the compiler wraps each MMA in an elect / R2UR.BROADCAST
    // waterfall of ~17 dependent instructions (measured: ~140 clk per MMA issued).  One elected lane issues.
    const int zo = __shfl_sync(0xffffffffu, warp - TC_MMA_WARP, 0);
    if (zo < TZ) {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(leader));
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      // D = f32, A = B = tf32, both K-major, N = NB, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
      // descriptors are built once; only their 14-bit address fields advance (everything sits below 256 KB)
      const uint64_t a0 = tc_desc(tc_smem_u32(s_stage) + zo * ROW_BYTES, 2048, 128), b0 = tc_desc(tc_smem_u32(s_w), NB * 16, 128);
      const bool three = p.npass == 3;
      uint32_t st = 0, ph = 0, ai = 0;
      tc_mbar_wait(wready, 0, 700);  // the staged weights (generic-proxy stores, fenced by their writers)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int item = cta; item < p.items_per_cot; item += p.ctas_per_cot) {
        for (int step = 0; step < nsteps; ++step) {
          const uint32_t ab = ai & 1, aph = (ai >> 1) & 1;
          tc_mbar_wait(&acce[ab], aph ^ 1, 400 + (int)ab);  // epilogue has drained this accumulator buffer
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d = tmem_u + ab * ACC_COLS + zo * NB;
          for (int cgb = 0; cgb < ncg; cgb += CGS) {
            tc_mbar_wait(&full[st], ph, 500 + (int)st);
#ifndef TC_FENCE_WRITER
            // generic-proxy stores of the producers (ordered before this point by their release / this acquire) ->
            // async proxy: the fence lies on the causality path between the stores and the MMAs below; this warp has
            // no loads in flight, so its MEMBAR is free
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t a_st = a0 + (uint64_t)((st * STAGE_BYTES) >> 4);
            if (leader) {
#pragma unroll
              for (int cgl = 0; cgl < CGS; ++cgl) {
                const int cg = cgb + cgl;
                if (cg < ncg) {
                  const uint64_t b_cg = b0 + (uint64_t)(((uint32_t)cg * KD * 2 * WSLAB) >> 4);
                  const uint32_t acc0 = cg > 0 ? 1u : 0u;
#pragma unroll
                  for (int kd = 0; kd < KD; ++kd) {
                    const uint64_t a_hi = a_st + (uint64_t)(((cgl * NROW + kd) * ROW_BYTES) >> 4);
                    const uint64_t b_hi = b_cg + (uint64_t)((kd * 2 * WSLAB) >> 4);
                    tc_mma(d, a_hi, b_hi, idesc, kd > 0 ? 1u : acc0);
                    if (three) {
                      tc_mma(d, a_hi + (4096 >> 4), b_hi, idesc, 1u);
                      tc_mma(d, a_hi, b_hi + (WSLAB >> 4), idesc, 1u);
                    }
                  }
                }
              }
              tc_commit(&empty[st]);  // the stage is free once the MMAs of every issuer have read it
            }
            __syncwarp();
            if (++st == (uint32_t)NS) {
              st = 0;
              ph ^= 1;
            }
          }
          if (leader) tc_commit(&accf[ab]);  // this plane's accumulator row of the y step is complete
          __syncwarp();
          ++ai;
        }
      }
    }
  } else {
    // ============================ epilogue ============================
    // warps w, w+4, ... share TMEM lane quadrant q = w % 4 (one strip) and own CW = COT / (NEW/4) channels each.  A
    // (z_o, 4 channels) unit goes TMEM -> 9-tap gather -> BN/activation -> global before the next one is read, so
    // only the rolling window stays live in registers; addresses are 32-bit offsets from one pointer per item.
    const int q = warp & 3;
    const int ch0 = (warp >> 2) * CW;  // first channel (within the tile) of this warp
    const int nvalid = p.Cout - (cot * COT + ch0);  // channels of this warp that exist
    const int oC = (int)p.oC, oD = (int)p.oD, oH = (int)p.oH;
    const bool post = p.out_mul || p.residual || p.act2 != ESM_ACT_NONE;
    const bool gelu = p.act == ESM_ACT_GELU;
    const float oscale = p.out_scale;
    // expected-value correction of the accumulator truncation (tc_common.cuh): every accumulator column chained
    // ncg * KD MMA triples
    const float debias = 1.0f + TC_TRUNC_BIAS * (float)(ncg * KD * (p.npass == 3 ? 3 : 1));
    uint32_t ai = 0;
    for (int item = cta; item < p.items_per_cot; item += p.ctas_per_cot) {
      const TcItem ti = tc_decode(p, item, TZ);
      const int strip = ti.grp * 4 + q;
      const int seg = strip % p.nseg, ys = strip / p.nseg;
      const int x = seg * p.segw + lane - HALO;
      const int ya = ys * p.rows;
      const int yb = min(ya + p.rows, p.H);
      const bool lane_ok = strip < p.nstrips && lane >= HALO && lane < p.segw + HALO && x < p.W;
      // first channel of this warp, plane z0, row 0, column x (host checks that the offsets below fit 32 bits)
      float* op = p.out + ((long long)ti.b * p.oB + (long long)(cot * COT + ch0) * p.oC + (long long)ti.z0 * p.oD + x);
      const float* rp = p.residual ? p.residual + ((long long)ti.b * p.oB + (long long)(cot * COT + ch0) * p.oC + (long long)ti.z0 * p.oD + x) : nullptr;
      const float* mp = p.out_mul ? p.out_mul + ((long long)ti.b * p.omB + (long long)(cot * COT + ch0) * p.omC + x) : nullptr;
      const int nz = min(TZ, p.D - ti.z0);  // planes of this item that exist
      constexpr int PW = TAPS == 9 ? CW : 1;
      float Pa[TZ][PW], Pb[TZ][PW];  // partial sums of output rows y_in-1 and y_in (k3 only)
#pragma unroll
      for (int zo = 0; zo < TZ; ++zo)
#pragma unroll
        for (int c = 0; c < PW; ++c) Pa[zo][c] = Pb[zo][c] = 0.f;
      for (int step = 0; step < nsteps; ++step) {
        const int yo = ya - 2 * HALO + step;  // the output row this step completes
        const bool row_ok = lane_ok && yo >= ya && yo < yb;
        const uint32_t ab = ai & 1, aph = (ai >> 1) & 1;
        tc_mbar_wait(&accf[ab], aph, 600 + (int)ab);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int zo = 0; zo < TZ; ++zo) {
          const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + ab * ACC_COLS + zo * NB + ch0 * TAPS;
#pragma unroll
          for (int c4 = 0; c4 < CW; c4 += 4) {
            float rv[4];
            if (TAPS == 1) {
              tc_ld4(tb + c4, rv);
              tc_ld_wait();
            } else {
              float d[36];  // [channel j][kh][kw]: 4 channels x 9 taps are 36 adjacent columns
              tc_ld16(tb + c4 * 9, d);
              tc_ld16(tb + c4 * 9 + 16, d + 16);
              tc_ld4(tb + c4 * 9 + 32, d + 32);
              tc_ld_wait();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // out column x gathers input columns x-1 (kw=0), x (kw=1), x+1 (kw=2); kh = 0/1/2 feeds output rows
                // y_in+1 / y_in / y_in-1 (the last one is then complete)
                const float t0 = __shfl_up_sync(0xffffffffu, d[j * 9 + 0], 1) + d[j * 9 + 1] + __shfl_down_sync(0xffffffffu, d[j * 9 + 2], 1);
                const float t1 = __shfl_up_sync(0xffffffffu, d[j * 9 + 3], 1) + d[j * 9 + 4] + __shfl_down_sync(0xffffffffu, d[j * 9 + 5], 1);
                const float t2 = __shfl_up_sync(0xffffffffu, d[j * 9 + 6], 1) + d[j * 9 + 7] + __shfl_down_sync(0xffffffffu, d[j * 9 + 8], 1);
                rv[j] = Pa[zo][c4 + j] + t2;
                Pa[zo][c4 + j] = Pb[zo][c4 + j] + t1;
                Pb[zo][c4 + j] = t0;
              }
            }
            if (zo == TZ - 1 && c4 + 4 >= CW) {  // last TMEM read of this step: hand the accumulator buffer back
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) tc_mbar_arrive(&acce[ab]);
            }
            if (row_ok && zo < nz) {
              const int cl = ch0 + c4;  // channel within the tile
#pragma unroll
              for (int j = 0; j < 4; ++j) rv[j] = fmaf(rv[j] * debias, s_aff[cl + j], s_aff[COT + cl + j]);
              if (gelu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rv[j] = tc_gelu(rv[j]);
              } else if (p.act != ESM_ACT_NONE) {
                const float4 r = apply_act4(make_float4(rv[0], rv[1], rv[2], rv[3]), p.act);
                rv[0] = r.x; rv[1] = r.y; rv[2] = r.z; rv[3] = r.w;
              }
              if (TAPS == 1 && p.ps == 2) {
                // UpShuffle (ESMStereo.py:265-268): the 4 channels of a unit are the 2 x 2 output pixels of shuffled
                // channel (co / 4) at (2y, 2x): two 8-byte stores per lane, 256 contiguous bytes per warp and row
                if (p.act2 == ESM_ACT_SILU) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) rv[j] = tc_silu(rv[j]);
                } else if (p.act2 != ESM_ACT_NONE) {
                  const float4 r = apply_act4(make_float4(rv[0], rv[1], rv[2], rv[3]), p.act2);
                  rv[0] = r.x; rv[1] = r.y; rv[2] = r.z; rv[3] = r.w;
                }
                if (c4 < nvalid) {
                  float* o = p.out + ((long long)ti.b * p.oB + (long long)((cot * COT + cl) >> 2) * p.oC + (long long)(2 * yo) * p.oH + 2 * x);
                  *reinterpret_cast<float2*>(o) = make_float2(rv[0] * oscale, rv[1] * oscale);
                  *reinterpret_cast<float2*>(o + p.oH) = make_float2(rv[2] * oscale, rv[3] * oscale);
                }
                continue;
              }
              const int o_off = zo * oD + yo * oH + c4 * oC;
              if (post) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (c4 + j < nvalid) {
                    if (mp) rv[j] *= __ldg(mp + ((c4 + j) * (int)p.omC + yo * (int)p.omH));
                    if (rp) rv[j] += __ldg(rp + (o_off + j * oC));
                  }
                }
                if (p.act2 != ESM_ACT_NONE) {
                  const float4 r = apply_act4(make_float4(rv[0], rv[1], rv[2], rv[3]), p.act2);
                  rv[0] = r.x; rv[1] = r.y; rv[2] = r.z; rv[3] = r.w;
                }
              }
              float* o = op + o_off;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (c4 + j < nvalid) o[j * oC] = rv[j] * oscale;
            }
          }
        }
        ++ai;
      }
    }
  }
#ifdef TC_PROFILE
  if (blockIdx.x == 0 && lane == 0) {
    if (warp == 0) printf("tc_prof prologue %lld clk, grid %d\n", tc_t1 - tc_t0, gridDim.x);
    printf("tc_prof warp %2d: done at %8lld clk; waits: empty %8llu  acce %8llu  full %8llu  accf %8llu\n", warp, clock64() - tc_t0,
           tc_prof_wait[warp][2], tc_prof_wait[warp][4], tc_prof_wait[warp][5], tc_prof_wait[warp][6]);
    for (int i = 0; i < 8; ++i) tc_prof_wait[warp][i] = 0;
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// kh in K, kw in N (2D k3 s1 p1, COT = 32 output channels per CTA).  The taps-in-N kernel above reads 9 accumulator
// columns per output from tensor memory (64 B/clk/SM: 2.3 k clk per 128 x 32 outputs against 0.9 k clk of MMAs), cannot
// hold more than 24 channels per CTA (9 * COT <= 256 columns), so a 32-channel layer runs as two channel tiles whose CTAs
// both load and split the same activations, and its epilogue carries the kh gather in a rolling register window.  Here
//     acc[y_out][x_in, (co, kw)] = sum_{kh, ci} X[y_out + kh - 1, x_in, ci] * W[kh, kw, ci, co]
// an input row tile (the same ring stage as above) is multiplied by the three kh weight slabs into the accumulators
// of the three output rows it touches (N = 3 * COT = 96), a ring of four accumulators in tensor memory; the epilogue
// reads 3 columns per output and only the kw gather (two warp shuffles) is left.  Same producers, same split-TF32
// arithmetic; an accumulator chains 9 * ncg MMAs (the epilogue's expected-value correction uses that count).
template <int COT, bool TS>
__global__ void __launch_bounds__(tc_threads(8), 1) tck_conv_kernel(const __grid_constant__ TcK p) {
  constexpr int TC_NEW = 8, TC_MMA_WARP = TC_NEW, TC_PROD_WARP = TC_NEW + tc_nmw(8);
  constexpr int N3 = 3 * COT;                  // accumulator columns per output row: (co, kw)
  // TS: the A operand (activation row tiles) lives in TENSOR memory and the MMAs are TS-form (tc_mma_ts): in shared
  // memory every M128 x N96 x K8 MMA read 4 KB of A next to 3 KB of B at 128 B/clk -- 83 clk per MMA against 56 in
  // isolation, the kernel's bound once the weights came by bulk copy.  Tensor memory: four accumulator slots of N3 = 96
  // columns, then a two-stage A ring of 4 groups x (hi 8 | lo 8) columns; the producers write their rows with
  // tcgen05.st (a warp reaches only its own lane quadrant: quadrant = warp % 4), no proxy fence, no shared-memory stage.
  constexpr int SLOT = TS ? N3 : 128, NSLOT = 4;  // accumulator ring in tensor memory
  constexpr uint32_t ACOL = NSLOT * N3;           // TS: first column of the A ring (two stages x 4 groups x 16 columns)
  static_assert(!TS || ACOL + 2 * 4 * 16 <= 512, "tensor memory budget");
  constexpr int CGS = 4, TPW = CGS;            // 8-channel groups (= row tiles) per stage
  constexpr int ROW_BYTES = 8192;              // row tile: hi [2][128][4] floats, then lo
  constexpr int STAGE_BYTES = TPW * ROW_BYTES;
  constexpr int WSLAB = N3 * 32;               // one (cg, kh, hi|lo) B operand: [2][N3][4] floats
  constexpr int CW = COT / 2;                  // output channels per epilogue warp
  static_assert(N3 <= SLOT && N3 % 8 == 0 && CW % 8 == 0, "accumulator slot");
  constexpr int STAGE_SMEM = TS ? 0 : STAGE_BYTES;  // shared-memory bytes per ring stage
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef TC_PROFILE
  const long long tc_t0 = clock64();
#endif
  const int ncg = p.ncg;
  const int NS = p.nstages;
  const uint32_t wbytes = (uint32_t)ncg * 3 * 2 * WSLAB;
  uint8_t* s_w = smem;
  uint8_t* s_stage = smem + ((wbytes + 127u) & ~127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_stage + (size_t)NS * STAGE_SMEM);
  uint64_t* empty = full + NS;
  uint64_t* accf = empty + NS;
  uint64_t* acce = accf + NSLOT;
  uint64_t* wready = acce + NSLOT;
  uint64_t* accl = wready + 1;  // once-only "accumulator full" of the CTA's last rows (helper warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accl + NSLOT);
  float* s_aff = reinterpret_cast<float*>(tmem_slot + 2);  // [2][COT] scale, shift of this channel tile
  const int cot = blockIdx.x / p.ctas_per_cot;
  const int cta = blockIdx.x % p.ctas_per_cot;
  const int rows = p.rows, nsteps = rows + 2;
  const bool HELP = p.tail_help != 0;  // producer warps take half of the last rows' epilogue

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      tc_mbar_init(&full[i], TC_NTW);
      tc_mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < NSLOT; ++i) {
      tc_mbar_init(&accf[i], 1);
      tc_mbar_init(&acce[i], TC_NEW);
      tc_mbar_init(&accl[i], 1);
    }
    tc_mbar_init(wready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 2 * COT) {
    const int c = tid % COT, co = cot * COT + c;
    const float* src = tid < COT ? p.scale : p.shift;
    s_aff[tid] = (src && co < p.Cout) ? __ldg(src + co) : (tid < COT ? 1.f : 0.f);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // resident weights: the pack holds this channel tile's image (split, UMMA layout: conv_tc.cuh TcImg), one bulk copy
  // per 8-channel group, all completing on `wready`; the MMA issuer waits for it before its first MMA
  if (tid == TC_MMA_WARP * 32) {
    tc_mbar_expect_tx(wready, wbytes);
    const float* img = p.wimg + (long long)cot * ncg * (3 * 2 * WSLAB / 4);
    for (int cg = 0; cg < ncg; ++cg) tc_bulk_g2s(s_w + (size_t)cg * (3 * 2 * WSLAB), img + (long long)cg * (3 * 2 * WSLAB / 4), 3 * 2 * WSLAB, wready);
  }

  // ---- epilogue of one output row of an item for TMEM lane quadrant q: channels [ch0 + cb, ch0 + ce) of the tile, one
  // 8-channel group (24 accumulator columns) per TMEM round trip: kw gather by two warp shuffles, BN / activation, stores.
  // `once`: wait on the once-only barrier of one of the CTA's last four rows (the helper warps below may be several
  // phases ahead of accf, and a parity can only name the current phase or the one before it).
  const int oC = (int)p.oC, oH = (int)p.oH;
  const bool post = p.out_mul || p.residual || p.act2 != ESM_ACT_NONE;
  const bool gelu = p.act == ESM_ACT_GELU;
  const float oscale = p.out_scale;
  const float debias = 1.0f + TC_TRUNC_BIAS * (float)(ncg * 3 * (p.npass == 3 ? 3 : 1));
  const int n_my = cta < p.items_per_cot ? (p.items_per_cot - cta + p.ctas_per_cot - 1) / p.ctas_per_cot : 0;  // items of this CTA
  const int tail_rows = rows < NSLOT ? rows : NSLOT;  // rows of the last item whose slots are never reused
  auto epi_row = [&](int item, int r, uint32_t rc, int q, int ch0, int cb, int ce, bool once, bool hand_back) {
    const TcItem ti = tc_decode(p, item, 1);
    const int strip = ti.grp * 4 + q;
    const int seg = strip % p.nseg, ys = strip / p.nseg;
    const int x = seg * p.segw + lane - 1;
    const int ya = ys * rows;
    const int yb = min(ya + rows, p.H);
    const bool lane_ok = strip < p.nstrips && lane >= 1 && lane < p.segw + 1 && x < p.W;
    const int nvalid = p.Cout - (cot * COT + ch0);
    float* op = p.out + ((long long)ti.b * p.oB + (long long)(cot * COT + ch0) * p.oC + x);
    const float* rp = p.residual ? p.residual + ((long long)ti.b * p.oB + (long long)(cot * COT + ch0) * p.oC + x) : nullptr;
    const float* mp = p.out_mul ? p.out_mul + ((long long)ti.b * p.omB + (long long)(cot * COT + ch0) * p.omC + x) : nullptr;
    {
      {
        const int yo = ya + r;
        const bool row_ok = lane_ok && yo < yb;
        tc_mbar_wait(once ? &accl[rc & 3] : &accf[rc & 3], once ? 0u : ((rc >> 2) & 1), 600 + (int)(rc & 3));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (rc & 3) * SLOT + ch0 * 3;
        for (int c8 = cb; c8 < ce; c8 += 8) {
          // two 4-channel units per TMEM round trip: their gather / BN / activation chains are independent, which is
          // what hides the dependent-issue latency of 8 epilogue warps (one unit at a time: 950 clk per unit)
          float d[24], rv[8];  // [channel j][kw]
          tc_ld16(tb + c8 * 3, d);
          tc_ld8(tb + c8 * 3 + 16, d + 16);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)  // out column x gathers input columns x-1 (kw=0), x (kw=1), x+1 (kw=2)
            rv[j] = __shfl_up_sync(0xffffffffu, d[j * 3 + 0], 1) + d[j * 3 + 1] + __shfl_down_sync(0xffffffffu, d[j * 3 + 2], 1);
          if (hand_back && c8 + 8 >= ce) {  // last TMEM read of this row: hand the accumulator slot back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&acce[rc & 3]);
          }
          if (row_ok) {
            const int cl = ch0 + c8;
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = fmaf(rv[j] * debias, s_aff[cl + j], s_aff[COT + cl + j]);
            if (gelu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) rv[j] = tc_gelu(rv[j]);
            } else if (p.act != ESM_ACT_NONE) {
#pragma unroll
              for (int h = 0; h < 8; h += 4) {
                const float4 a = apply_act4(make_float4(rv[h], rv[h + 1], rv[h + 2], rv[h + 3]), p.act);
                rv[h] = a.x; rv[h + 1] = a.y; rv[h + 2] = a.z; rv[h + 3] = a.w;
              }
            }
            const int o_off = yo * oH + c8 * oC;
            if (post) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (c8 + j < nvalid) {
                  if (mp) rv[j] *= __ldg(mp + ((c8 + j) * (int)p.omC + yo * (int)p.omH));
                  if (rp) rv[j] += __ldg(rp + (o_off + j * oC));
                }
              }
              if (p.act2 != ESM_ACT_NONE) {
#pragma unroll
                for (int h = 0; h < 8; h += 4) {
                  const float4 a = apply_act4(make_float4(rv[h], rv[h + 1], rv[h + 2], rv[h + 3]), p.act2);
                  rv[h] = a.x; rv[h + 1] = a.y; rv[h + 2] = a.z; rv[h + 3] = a.w;
                }
              }
            }
            float* o = op + o_off;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c8 + j < nvalid) o[j * oC] = rv[j] * oscale;
          }
        }
      }
    }
  };

  if (warp >= TC_PROD_WARP) {
    // ============================ operand producers (as in tc_conv_kernel, 2D, one input row per step) ============================
    const int tw = warp - TC_PROD_WARP;
    const int q = TS ? (warp & 3) : ((tw >> 1) & 3);   // strip = TMEM lane quadrant (TS: the one this warp may write)
    const int khalf = TS ? (tw >> 2) : (tw & 1);        // which 4 of the 8 channels of a group
    const int m = q * 32 + lane;
    uint32_t st = 0, ph = 0;
    int item = cta, step = 0, cgb = 0;
    const float* base[3] = {nullptr, nullptr, nullptr};
    int ya = 0, x = 0;
    bool strip_ok = false;
    auto enter_item = [&]() {
      if (item >= p.items_per_cot) return;
      const TcItem ti = tc_decode(p, item, 1);
      const int strip = ti.grp * 4 + q;
      const int seg = strip % p.nseg, ys = strip / p.nseg;
      x = seg * p.segw + lane - 1;
      ya = ys * rows;
      strip_ok = strip < p.nstrips && x >= 0 && x < p.W && lane < p.segw + 2;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (i < p.nsrc) base[i] = p.src[i].ptr + (long long)ti.b * p.src[i].sB;
    };
    auto load = [&](float (&v)[TPW][4]) {
      const int y = ya - 1 + step;
      const bool ok = strip_ok && (unsigned)y < (unsigned)p.H;
#pragma unroll
      for (int cgl = 0; cgl < CGS; ++cgl) {
        int rel = (cgb + cgl) * 8 + khalf * 4, k = 0;
        if (p.nsrc > 1) {
          while (k < p.nsrc - 1 && rel >= p.src[k].C) {  // host guarantees 8-channel groups never straddle sources
            rel -= p.src[k].C;
            ++k;
          }
        }
        const int sC = (int)p.src[k].sC;
        const int nch = p.src[k].C - rel;
        const float* bp = p.nsrc > 1 ? (k == 0 ? base[0] : k == 1 ? base[1] : base[2]) : base[0];
        const int off = rel * sC + y * (int)p.src[k].sH + x;
#pragma unroll
        for (int c = 0; c < 4; ++c) v[cgl][c] = (ok && c < nch) ? __ldg(bp + (off + c * sC)) : 0.f;
      }
    };
    auto advance = [&]() {
      cgb += CGS;
      if (cgb >= ncg) {
        cgb = 0;
        if (++step >= nsteps) {
          step = 0;
          item += p.ctas_per_cot;
          enter_item();
        }
      }
    };
    auto store_stage = [&](const float (&v)[TPW][4]) {
      tc_mbar_wait(&empty[st], ph ^ 1, 200 + (int)st);
      if constexpr (TS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + ACOL + st * (TPW * 16) + khalf * 4;
#pragma unroll
        for (int j = 0; j < TPW; ++j) {
          const float h0 = tc_rna(v[j][0]), h1 = tc_rna(v[j][1]), h2 = tc_rna(v[j][2]), h3 = tc_rna(v[j][3]);
          tc_st4(ta + j * 16, h0, h1, h2, h3);
          if (p.npass == 3) tc_st4(ta + j * 16 + 8, tc_lo(v[j][0], h0), tc_lo(v[j][1], h1), tc_lo(v[j][2], h2), tc_lo(v[j][3], h3));
        }
        tc_st_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(&full[st]);
        if (++st == (uint32_t)NS) {
          st = 0;
          ph ^= 1;
        }
        return;
      }
      uint8_t* sb = s_stage + (size_t)st * STAGE_BYTES + khalf * 2048 + m * 16;
#pragma unroll
      for (int j = 0; j < TPW; ++j) {
        float4 hi, lo;
        hi.x = tc_rna(v[j][0]);
        hi.y = tc_rna(v[j][1]);
        hi.z = tc_rna(v[j][2]);
        hi.w = tc_rna(v[j][3]);
        *reinterpret_cast<float4*>(sb + j * ROW_BYTES) = hi;
        if (p.npass == 3) {
          lo.x = tc_lo(v[j][0], hi.x);
          lo.y = tc_lo(v[j][1], hi.y);
          lo.z = tc_lo(v[j][2], hi.z);
          lo.w = tc_lo(v[j][3], hi.w);
          *reinterpret_cast<float4*>(sb + j * ROW_BYTES + 4096) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&full[st]);
      if (++st == (uint32_t)NS) {
        st = 0;
        ph ^= 1;
      }
    };
    float va[TPW][4], vb[TPW][4];
    enter_item();
    if (item < p.items_per_cot) load(va);
    while (item < p.items_per_cot) {
      advance();
      if (item < p.items_per_cot) load(vb);
      store_stage(va);
      if (item >= p.items_per_cot) break;
      advance();
      if (item < p.items_per_cot) load(va);
      store_stage(vb);
    }
    // nothing left to load: take the second 8 channels of epilogue warp (warp % 4, tw / 4) on the last rows
    if (HELP && n_my > 0) {
      const int last_item = cta + (n_my - 1) * p.ctas_per_cot;
      const uint32_t rowc = (uint32_t)(n_my - 1) * (uint32_t)rows;
      for (int r = rows - tail_rows; r < rows; ++r) epi_row(last_item, r, rowc + (uint32_t)r, warp & 3, (tw >> 2) * CW, 8, CW, true, false);
    }
  } else if (warp == TC_MMA_WARP) {
    // ============================ MMA issuer ============================
    // Step s brings input row ya - 1 + s: kh = 0 / 1 / 2 send it to output rows s / s-1 / s-2 of the item (where they
    // exist).  Output row r is first written at step r (kh = 0, first group: that MMA overwrites) after the epilogue
    // has drained the slot's previous row, and is complete after step r + 2.  The whole warp walks the loop so that the
    // descriptors stay in uniform registers; one elected lane issues.
    const uint32_t leader = tc_elect();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N3 >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = tc_desc(tc_smem_u32(s_stage), 2048, 128), b0 = tc_desc(tc_smem_u32(s_w), N3 * 16, 128);
    const bool three = p.npass == 3;
    uint32_t st = 0, ph = 0, rowc = 0;  // rowc: output rows of the items before this one
    tc_mbar_wait(wready, 0, 700);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int it = 0;
    for (int item = cta; item < p.items_per_cot; item += p.ctas_per_cot, ++it) {
      for (int step = 0; step < nsteps; ++step) {
        if (step < rows) {
          const uint32_t rc = rowc + (uint32_t)step;
          tc_mbar_wait(&acce[rc & 3], ((rc >> 2) & 1) ^ 1, 400 + (int)(rc & 3));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int cgb = 0; cgb < ncg; cgb += CGS) {
          tc_mbar_wait(&full[st], ph, 500 + (int)st);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t a_st = a0 + (uint64_t)((st * STAGE_BYTES) >> 4);
          if (leader) {
#pragma unroll
            for (int cgl = 0; cgl < CGS; ++cgl) {
              const int cg = cgb + cgl;
              if (cg < ncg) {
                const uint64_t a_hi = a_st + (uint64_t)((cgl * ROW_BYTES) >> 4);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                  const int r = step - kh;
                  if (r >= 0 && r < rows) {
                    const uint32_t d = tmem_u + ((rowc + (uint32_t)r) & 3) * SLOT;
                    const uint64_t b_hi = b0 + (uint64_t)(((uint32_t)(cg * 3 + kh) * 2 * WSLAB) >> 4);
                    if constexpr (TS) {
                      const uint32_t ta = tmem_u + ACOL + (st * CGS + cgl) * 16;  // hi k0..7 | lo k0..7
                      tc_mma_ts(d, ta, b_hi, idesc, (kh == 0 && cg == 0) ? 0u : 1u);
                      if (three) {
                        tc_mma_ts(d, ta + 8, b_hi, idesc, 1u);
                        tc_mma_ts(d, ta, b_hi + (WSLAB >> 4), idesc, 1u);
                      }
                    } else {
                      tc_mma(d, a_hi, b_hi, idesc, (kh == 0 && cg == 0) ? 0u : 1u);
                      if (three) {
                        tc_mma(d, a_hi + (4096 >> 4), b_hi, idesc, 1u);
                        tc_mma(d, a_hi, b_hi + (WSLAB >> 4), idesc, 1u);
                      }
                    }
                  }
                }
              }
            }
            tc_commit(&empty[st]);
          }
          __syncwarp();
          if (++st == (uint32_t)NS) {
            st = 0;
            ph ^= 1;
          }
        }
        if (step >= 2) {
          if (leader) {
            tc_commit(&accf[(rowc + (uint32_t)(step - 2)) & 3]);
            if (HELP && it == n_my - 1 && step - 2 >= rows - tail_rows) tc_commit(&accl[(rowc + (uint32_t)(step - 2)) & 3]);
          }
          __syncwarp();
        }
      }
      rowc += (uint32_t)rows;
    }
  } else if (warp < TC_NEW) {
    // ============================ epilogue ============================
    // warp w and w+4 share TMEM lane quadrant q = w % 4 (one strip) and own CW = COT/2 channels each.  On the last
    // tail_rows rows of the CTA's last item each of them keeps the first 8 of its 16 channels and a producer warp, idle
    // by then, takes the other 8 (below): those rows' epilogue is exposed (nothing overlaps it).
    const int q = warp & 3;
    const int ch0 = (warp >> 2) * CW;
    uint32_t rowc = 0;
    int i = 0;
    for (int item = cta; item < p.items_per_cot; item += p.ctas_per_cot, ++i) {
      for (int r = 0; r < rows; ++r) {
        const bool tail = HELP && i == n_my - 1 && r >= rows - tail_rows;
        epi_row(item, r, rowc + (uint32_t)r, q, ch0, 0, tail ? 8 : CW, false, true);
      }
      rowc += (uint32_t)rows;
    }
  }
#ifdef TC_PROFILE
  if (blockIdx.x == 0 && lane == 0) {
    printf("tck_prof warp %2d: done at %8lld clk (grid %d rows %d); waits: empty %8llu  acce %8llu  full %8llu  accf %8llu  wready %8llu\n", warp, clock64() - tc_t0,
           gridDim.x, rows, tc_prof_wait[warp][2], tc_prof_wait[warp][4], tc_prof_wait[warp][5], tc_prof_wait[warp][6], tc_prof_wait[warp][7]);
    for (int i = 0; i < 8; ++i) tc_prof_wait[warp][i] = 0;
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Pointwise (k1) layers, TS-form: tc1_conv_kernel.  The taps-in-N kernel runs k1 layers as TAPS = 1 over image strips
// (32 columns of one row per TMEM lane quadrant) with its shared-memory operand ring, whose proxy fence (MEMBAR) makes a
// producer stage cost one exposed load latency on top of its stores: these layers -- 100-160 input channels, 16-64
// outputs, 0.2 GFLOP -- are bound by exactly that.  Here a tile is 128 CONSECUTIVE pixels of the flattened (d, h, w)
// lattice (no strips, no idle lanes on narrow images), the producers write their rows of the activation tile straight
// into a six-stage ring in tensor memory (tcgen05.st, no fence: the loads of the next stage stay in flight under the
// stores of this one), the MMAs are TS-form with the resident pre-split weight image as B, and the accumulator
// (N = COT <= 64 columns) is double-buffered.  Same split-TF32 arithmetic, same epilogue options (BN, activations,
// broadcast multiply, residual, second activation, scale) as the taps-in-N kernel; PixelShuffle stays there.
template <int COT>
__global__ void __launch_bounds__(tc_threads(8), 1) tc1_conv_kernel(const __grid_constant__ TcK p) {
  constexpr int TC_NEW = 8, TC_MMA_WARP = TC_NEW, TC_PROD_WARP = TC_NEW + tc_nmw(8);
  constexpr int CGS = 4;                        // 8-channel groups per ring stage
  constexpr int ACC = 64;                       // columns per accumulator buffer (COT <= 64), two buffers
  constexpr uint32_t ACOL = 2 * ACC;            // first column of the A ring
  constexpr int STAGE_COLS = CGS * 16;          // hi 8 | lo 8 columns per group
  constexpr int NSA = (512 - (int)ACOL) / STAGE_COLS;  // 6 ring stages
  constexpr int WSLAB = COT * 32;               // bytes of one (cg, hi|lo) B operand: [2][COT][4] floats
  constexpr int CW = COT / 2;                   // output channels per epilogue warp
  static_assert(COT % 8 == 0 && COT <= ACC && CW % 4 == 0, "channel tile");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncg = p.ncg;
  const uint32_t wbytes = (uint32_t)ncg * 2 * WSLAB;
  uint8_t* s_w = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ((wbytes + 127u) & ~127u));
  uint64_t* empty = full + NSA;
  uint64_t* accf = empty + NSA;
  uint64_t* acce = accf + 2;
  uint64_t* wready = acce + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wready + 1);
  float* s_aff = reinterpret_cast<float*>(tmem_slot + 2);  // [2][COT] scale, shift
  const int npix = p.D * p.H * p.W;             // pixels of one batch item
  const int tiles = (npix + 127) >> 7;
  const int items = p.B * tiles;

  if (tid == 0) {
    for (int i = 0; i < NSA; ++i) {
      tc_mbar_init(&full[i], TC_NTW);
      tc_mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc_mbar_init(&accf[i], 1);
      tc_mbar_init(&acce[i], TC_NEW);
    }
    tc_mbar_init(wready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 2 * COT) {
    const int c = tid % COT;
    const float* src = tid < COT ? p.scale : p.shift;
    s_aff[tid] = (src && c < p.Cout) ? __ldg(src + c) : (tid < COT ? 1.f : 0.f);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (tid == TC_MMA_WARP * 32) {  // resident weights: the pack's pre-split image, one bulk copy per 8-channel group
    tc_mbar_expect_tx(wready, wbytes);
    for (int cg = 0; cg < ncg; ++cg) tc_bulk_g2s(s_w + (size_t)cg * (2 * WSLAB), p.wimg + (long long)cg * (2 * WSLAB / 4), 2 * WSLAB, wready);
  }

  if (warp >= TC_PROD_WARP) {
    // ============================ A-operand producers ============================
    const int tw = warp - TC_PROD_WARP;
    const int q = warp & 3;        // the TMEM lane quadrant this warp may write = rows q*32 .. q*32+31 of the tile
    const int khalf = tw >> 2;     // which 4 of the 8 channels of a group
    uint32_t st = 0, ph = 0;
    int item = blockIdx.x, cgb = 0;
    const float* base[3] = {nullptr, nullptr, nullptr};  // this lane's pixel in each source
    bool ok = false;
    auto enter_item = [&]() {
      if (item >= items) return;
      const int b = item / tiles, pix = (item - b * tiles) * 128 + q * 32 + lane;
      ok = pix < npix;
      const int x = pix % p.W, r = pix / p.W;
      const int y = r % p.H, z = r / p.H;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (i < p.nsrc) base[i] = p.src[i].ptr + ((long long)b * p.src[i].sB + (long long)z * p.src[i].sD + (long long)y * p.src[i].sH + x);
    };
    auto load = [&](float (&v)[CGS][4]) {
#pragma unroll
      for (int cgl = 0; cgl < CGS; ++cgl) {
        int rel = (cgb + cgl) * 8 + khalf * 4, k = 0;
        if (p.nsrc > 1) {
          while (k < p.nsrc - 1 && rel >= p.src[k].C) {  // host guarantees 8-channel groups never straddle sources
            rel -= p.src[k].C;
            ++k;
          }
        }
        const int sC = (int)p.src[k].sC;
        const int nch = p.src[k].C - rel;  // valid channels from `rel` on (<= 0 past the last group)
        const float* bp = p.nsrc > 1 ? (k == 0 ? base[0] : k == 1 ? base[1] : base[2]) : base[0];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[cgl][c] = (ok && c < nch) ? __ldg(bp + (rel + c) * sC) : 0.f;
      }
    };
    auto advance = [&]() {
      cgb += CGS;
      if (cgb >= ncg) {
        cgb = 0;
        item += gridDim.x;
        enter_item();
      }
    };
    auto store_stage = [&](const float (&v)[CGS][4]) {
      tc_mbar_wait(&empty[st], ph ^ 1, 200 + (int)st);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + ACOL + st * STAGE_COLS + khalf * 4;
#pragma unroll
      for (int j = 0; j < CGS; ++j) {
        const float h0 = tc_rna(v[j][0]), h1 = tc_rna(v[j][1]), h2 = tc_rna(v[j][2]), h3 = tc_rna(v[j][3]);
        tc_st4(ta + j * 16, h0, h1, h2, h3);
        if (p.npass == 3) tc_st4(ta + j * 16 + 8, tc_lo(v[j][0], h0), tc_lo(v[j][1], h1), tc_lo(v[j][2], h2), tc_lo(v[j][3], h3));
      }
      tc_st_wait();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&full[st]);
      if (++st == (uint32_t)NSA) {
        st = 0;
        ph ^= 1;
      }
    };
    float va[CGS][4], vb[CGS][4];
    enter_item();
    if (item < items) load(va);
    while (item < items) {
      advance();
      if (item < items) load(vb);
      store_stage(va);
      if (item >= items) break;
      advance();
      if (item < items) load(va);
      store_stage(vb);
    }
  } else if (warp == TC_MMA_WARP) {
    // ============================ MMA issuer ============================
    const uint32_t leader = tc_elect();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(COT >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t b0 = tc_desc(tc_smem_u32(s_w), COT * 16, 128);
    const bool three = p.npass == 3;
    uint32_t st = 0, ph = 0, ai = 0;
    tc_mbar_wait(wready, 0, 700);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const uint32_t ab = ai & 1, aph = (ai >> 1) & 1;
      tc_mbar_wait(&acce[ab], aph ^ 1, 400 + (int)ab);  // the epilogue has drained this accumulator buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d = tmem_u + ab * ACC;
      for (int cgb = 0; cgb < ncg; cgb += CGS) {
        tc_mbar_wait(&full[st], ph, 500 + (int)st);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_st = tmem_u + ACOL + st * STAGE_COLS;
        const uint64_t b_st = b0 + (uint64_t)(((uint32_t)cgb * 2 * WSLAB) >> 4);
        if (leader) {
#pragma unroll
          for (int cgl = 0; cgl < CGS; ++cgl) {
            if (cgb + cgl < ncg) {
              const uint32_t a_hi = a_st + cgl * 16;
              const uint64_t b_hi = b_st + (uint64_t)((cgl * 2 * WSLAB) >> 4);
              tc_mma_ts(d, a_hi, b_hi, idesc, (cgb + cgl) > 0 ? 1u : 0u);
              if (three) {
                tc_mma_ts(d, a_hi + 8, b_hi, idesc, 1u);
                tc_mma_ts(d, a_hi, b_hi + (WSLAB >> 4), idesc, 1u);
              }
            }
          }
          tc_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == (uint32_t)NSA) {
          st = 0;
          ph ^= 1;
        }
      }
      if (leader) tc_commit(&accf[ab]);
      __syncwarp();
      ++ai;
    }
  } else if (warp < TC_NEW) {
    // ============================ epilogue ============================
    // warp w and w+4 share TMEM lane quadrant q = w % 4 and own CW = COT/2 channels each, 8 (or the last 4) at a time
    const int q = warp & 3;
    const int ch0 = (warp >> 2) * CW;
    const int nvalid = p.Cout - ch0;
    const int oC = (int)p.oC;
    const bool post = p.out_mul || p.residual || p.act2 != ESM_ACT_NONE;
    const bool gelu = p.act == ESM_ACT_GELU;
    const float oscale = p.out_scale;
    const float debias = 1.0f + TC_TRUNC_BIAS * (float)(ncg * (p.npass == 3 ? 3 : 1));
    uint32_t ai = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int b = item / tiles, pix = (item - b * tiles) * 128 + q * 32 + lane;
      const bool ok = pix < npix;
      const int x = pix % p.W, r = pix / p.W;
      const int y = r % p.H, z = r / p.H;
      const long long obase = (long long)b * p.oB + (long long)ch0 * p.oC + (long long)z * p.oD + (long long)y * p.oH + x;
      float* op = p.out + obase;
      const float* rp = p.residual ? p.residual + obase : nullptr;
      const float* mp = p.out_mul ? p.out_mul + ((long long)b * p.omB + (long long)ch0 * p.omC + (long long)y * p.omH + x) : nullptr;
      const uint32_t ab = ai & 1, aph = (ai >> 1) & 1;
      tc_mbar_wait(&accf[ab], aph, 600 + (int)ab);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + ab * ACC + ch0;
#pragma unroll
      for (int c8 = 0; c8 < CW; c8 += 8) {
        const int nc = (CW - c8) < 8 ? (CW - c8) : 8;  // 8, or 4 for the last unit of CW = 12 / 20
        float rv[8];
        if (CW - c8 >= 8) {
          tc_ld8(tb + c8, rv);
        } else {
          tc_ld4(tb + c8, rv);
          rv[4] = rv[5] = rv[6] = rv[7] = 0.f;
        }
        tc_ld_wait();
        if (c8 + 8 >= CW) {  // last TMEM read of this item: hand the accumulator buffer back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) tc_mbar_arrive(&acce[ab]);
        }
        if (ok) {
          const int cl = ch0 + c8;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < nc) rv[j] = fmaf(rv[j] * debias, s_aff[cl + j], s_aff[COT + cl + j]);
          if (gelu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = tc_gelu(rv[j]);
          } else if (p.act != ESM_ACT_NONE) {
#pragma unroll
            for (int h = 0; h < 8; h += 4) {
              const float4 a = apply_act4(make_float4(rv[h], rv[h + 1], rv[h + 2], rv[h + 3]), p.act);
              rv[h] = a.x; rv[h + 1] = a.y; rv[h + 2] = a.z; rv[h + 3] = a.w;
            }
          }
          if (post) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < nc && c8 + j < nvalid) {
                if (mp) rv[j] *= __ldg(mp + (long long)(c8 + j) * p.omC);
                if (rp) rv[j] += __ldg(rp + (long long)(c8 + j) * oC);
              }
            }
            if (p.act2 != ESM_ACT_NONE) {
#pragma unroll
              for (int h = 0; h < 8; h += 4) {
                const float4 a = apply_act4(make_float4(rv[h], rv[h + 1], rv[h + 2], rv[h + 3]), p.act2);
                rv[h] = a.x; rv[h + 1] = a.y; rv[h + 2] = a.z; rv[h + 3] = a.w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < nc && c8 + j < nvalid) op[(long long)(c8 + j) * oC] = rv[j] * oscale;
        }
      }
      ++ai;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

typedef void (*tc_fn_t)(const TcK);
static long long tc_launches = 0;

static tc_fn_t tc1_pick(int COT) {
  switch (COT) {
    case 8: return tc1_conv_kernel<8>;
    case 16: return tc1_conv_kernel<16>;
    case 24: return tc1_conv_kernel<24>;
    case 32: return tc1_conv_kernel<32>;
    case 40: return tc1_conv_kernel<40>;
    case 48: return tc1_conv_kernel<48>;
    case 64: return tc1_conv_kernel<64>;
    default: return nullptr;
  }
}

static tc_fn_t tc_pick(int COT, int TZ, int KD, bool gwc, int taps, int NEW) {
  if (NEW == 16) {  // four epilogue warps per quadrant: whole 4-channel units per warp, one output plane
    if (gwc || TZ != 1) return nullptr;
    if (taps == 9 && COT == 16) return KD == 1 ? (tc_fn_t)tc_conv_kernel<16, 1, 1, false, 9, 16> : KD == 3 ? (tc_fn_t)tc_conv_kernel<16, 1, 3, false, 9, 16> : nullptr;
    if (taps == 1 && KD == 1) {
      if (COT == 16) return tc_conv_kernel<16, 1, 1, false, 1, 16>;
      if (COT == 32) return tc_conv_kernel<32, 1, 1, false, 1, 16>;
      if (COT == 48) return tc_conv_kernel<48, 1, 1, false, 1, 16>;
      if (COT == 64) return tc_conv_kernel<64, 1, 1, false, 1, 16>;
    }
    return nullptr;
  }
  if (taps == 1) {  // pointwise: one channel tile holds the whole Cout (<= 64)
    if (gwc || TZ != 1 || KD != 1) return nullptr;
    switch (COT) {
      case 8: return tc_conv_kernel<8, 1, 1, false, 1>;
      case 16: return tc_conv_kernel<16, 1, 1, false, 1>;
      case 24: return tc_conv_kernel<24, 1, 1, false, 1>;
      case 32: return tc_conv_kernel<32, 1, 1, false, 1>;
      case 40: return tc_conv_kernel<40, 1, 1, false, 1>;
      case 48: return tc_conv_kernel<48, 1, 1, false, 1>;
      case 64: return tc_conv_kernel<64, 1, 1, false, 1>;
      default: return nullptr;
    }
  }
  if (gwc) return (COT == 8 && TZ == 3 && KD == 3) ? (tc_fn_t)tc_conv_kernel<8, 3, 3, true> : nullptr;
  if (KD == 3) {
    if (COT == 8 && TZ == 3) return tc_conv_kernel<8, 3, 3, false>;
    if (COT == 16 && TZ == 1) return tc_conv_kernel<16, 1, 3, false>;
    if (COT == 24 && TZ == 1) return tc_conv_kernel<24, 1, 3, false>;
  } else if (TZ == 1) {
    if (COT == 8) return tc_conv_kernel<8, 1, 1, false>;
    if (COT == 16) return tc_conv_kernel<16, 1, 1, false>;
    if (COT == 24) return tc_conv_kernel<24, 1, 1, false>;
  }
  return nullptr;
}

bool tc_conv_plan(const esm_conv_t* d, int num_sms, int npass, TcPlan* plan) {
  if (d->transposed || d->stride != 1) return false;
  const bool k3 = d->kh == 3 && d->kw == 3 && d->ph == 1 && d->pw == 1 && (d->kd == 1 || d->kd == 3) && d->pd == d->kd / 2;
  const bool k1 = d->kh == 1 && d->kw == 1 && d->kd == 1 && d->ph == 0 && d->pw == 0 && d->pd == 0;
  if (!k3 && !k1) return false;
  if (d->in_mul) return false;
  if (d->pixel_shuffle && !(d->pixel_shuffle == 2 && k1 && d->Cout % 4 == 0 && !d->out_mul && !d->residual && (d->oH % 2) == 0 &&
                            (reinterpret_cast<uintptr_t>(d->out) & 7) == 0))
    return false;
  if (d->Cin < 8 || num_sms <= 0) return false;
  // the kernel addresses one batch item with 32-bit element offsets
  for (int i = 0; i < d->nsrc; ++i)
    if (d->src[i].sB >= (1ll << 31) || (long long)d->src[i].C * d->src[i].sC >= (1ll << 31)) return false;
  if ((long long)d->Cout * d->oC >= (1ll << 31)) return false;
  const bool gwc = d->src_mode == ESM_SRC_GWC;
  if (gwc) {
    if (d->nsrc != 2 || d->src[0].C != 2 * d->Cin || d->kd != 3) return false;
  } else {
    for (int i = 0; i + 1 < d->nsrc; ++i)
      if (d->src[i].C % 8) return false;
  }
  const int CoutPad8 = round_up(d->Cout, 8);
  const int taps = k1 ? 1 : 9;
  int COT = tc_cot(d->Cout, k1);
  if (COT == 0) return false;
  if (gwc && COT != 8) return false;
  // 2D k3 layers whose Cout is a multiple of 32: kh in K, kw in N, 32 channels per CTA (tck_conv_kernel), when the
  // resident weights (18 KB per 8-channel group) leave room for two ring stages.  ESM_TC_KHK=0 keeps taps-in-N.
  static const bool khk_env = !(getenv("ESM_TC_KHK") && atoi(getenv("ESM_TC_KHK")) == 0);
  plan->khk = 0;
  if (khk_env && k3 && d->kd == 1 && !gwc && !d->pixel_shuffle && tc_img_geom(d->Cout, d->Cin, d->kd, d->kh, d->kw, 0).kind == 1 &&
      (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0) {
    // 2: A operand in tensor memory (TS-form MMAs; no shared-memory stages, so Cin <= 96 fits); 1 (ESM_TC_KHK_TS=0):
    // A in shared memory, which needs room for two 32 KB ring stages next to the weights (Cin <= 64)
    static const bool ts_env = !(getenv("ESM_TC_KHK_TS") && atoi(getenv("ESM_TC_KHK_TS")) == 0);
    const size_t wb = (size_t)ceil_div(d->Cin, 8) * 3 * 2 * 96 * 32 + 127;
    if (ts_env) {
      plan->khk = 2;
      COT = 32;
    } else if (wb + 2 * 4 * 8192 <= 227 * 1024 - 1024) {
      plan->khk = 1;
      COT = 32;
    }
  }
  // pointwise layers: the TS-form kernel over flat pixel tiles (tc1_conv_kernel) wherever the pack carries the image
  static const bool k1ts_env = !(getenv("ESM_TC_K1TS") && atoi(getenv("ESM_TC_K1TS")) == 0);
  plan->k1ts = 0;
  if (k1ts_env && k1 && !d->pixel_shuffle && (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0) {
    const TcImg tk = tc_img_geom(d->Cout, d->Cin, d->kd, d->kh, d->kw, 0);
    const size_t wb = (size_t)ceil_div(d->Cin, 8) * 2 * COT * 32;
    if (tk.kind == 2 && tk.COT == COT && tk.taps == 1 && tk.ncot == 1 && wb + 1024 <= 227 * 1024 && tc1_pick(COT)) plan->k1ts = 1;
  }
  plan->COT = COT;
  plan->taps = taps;
  plan->ncot = ceil_div(d->Cout, COT);
  plan->KD = d->kd;
  plan->TZ = (d->kd == 3 && COT == 8) ? 3 : 1;
  plan->gwc = gwc;
  plan->npass = npass;
  // ESM_TC_EPI=16: four epilogue warps per quadrant wherever a kernel exists for it.  Measured SLOWER (32->32 k3 at
  // 192x624: 29.0 against 27.2 us; 800 threads cap the kernel at 72 registers): default 8.
  static const int epi_env = getenv("ESM_TC_EPI") ? atoi(getenv("ESM_TC_EPI")) : 8;
  plan->NEW = (!plan->khk && epi_env == 16 && tc_pick(COT, plan->TZ, plan->KD, gwc, taps, 16)) ? 16 : 8;
  if (!plan->khk && !tc_pick(COT, plan->TZ, plan->KD, gwc, taps, plan->NEW)) return false;
  const int NB = taps * COT, NROW = plan->TZ + plan->KD - 1, CGS = NROW == 1 ? 4 : 1;
  const int ncg = ceil_div(d->Cin, 8);
  const size_t wbytes = plan->khk ? (((size_t)ncg * 3 * 2 * (3 * COT) * 32 + 127) & ~(size_t)127) : (((size_t)ncg * plan->KD * 2 * NB * 32 + 127) & ~(size_t)127);
  const size_t stage = (size_t)CGS * NROW * 8192;
  const size_t limit = 227 * 1024 - 1024;
  if (plan->k1ts) {
    const long long items = (long long)d->B * (((long long)d->Dout * d->Hout * d->Wout + 127) / 128);
    if (items >= (1ll << 30) || (long long)d->Dout * d->Hout * d->Wout >= (1ll << 31)) return false;
    plan->ncot = 1;
    plan->nstages = 6;
    plan->smem = (size_t)ceil_div(d->Cin, 8) * 2 * COT * 32 + 1024;
    plan->nseg = plan->segw = plan->ysplit = plan->rows = 1;
    plan->ctas_per_cot = (int)(items < num_sms ? items : num_sms);
    return true;
  }
  if (plan->khk == 2) {
    if (wbytes > limit) return false;
    plan->nstages = 2;  // the A ring lives in tensor memory
    plan->smem = wbytes + 1024;
  } else {
    if (wbytes + 2 * stage > limit) return false;
    int ns = (int)((limit - wbytes) / stage);
    plan->nstages = ns > 4 ? 4 : ns;
    plan->smem = wbytes + plan->nstages * stage + 1024;
  }
  const int segmax = k1 ? 32 : 30;
  plan->nseg = ceil_div(d->Wout, segmax);
  plan->segw = ceil_div(d->Wout, plan->nseg);
  const int ztiles = ceil_div(d->Dout, plan->TZ);
  const int sms = num_sms / plan->ncot > 0 ? num_sms / plan->ncot : 1;
  double best = 1e30;
  for (int ys = 1; ys <= 64 && ys <= d->Hout; ++ys) {
    const int rows = ceil_div(d->Hout, ys);
    if (ceil_div(d->Hout, rows) != ys) continue;
    const long long items = (long long)d->B * ztiles * ceil_div(plan->nseg * ys, 4);
    const long long waves = (items + sms - 1) / sms;
    const double cost = (double)waves * (rows + (k1 ? 0 : 2)) + 1.0;  // +1: per-item pipeline fill
    if (cost < best) {
      best = cost;
      plan->ysplit = ys;
      plan->rows = rows;
      plan->ctas_per_cot = (int)(items < sms ? items : sms);
    }
  }
  return true;
}

int tc_conv_launch(const esm_conv_t* d, const TcPlan& plan, cudaStream_t st) {
  TcK k;
  memset(&k, 0, sizeof(k));
  for (int i = 0; i < d->nsrc; ++i) k.src[i] = d->src[i];
  k.nsrc = d->nsrc;
  k.cpg = plan.gwc ? d->src[0].C / d->Cin : 0;
  k.B = d->B;
  k.Cin = d->Cin;
  k.ncg = ceil_div(d->Cin, 8);
  k.D = d->Dout;
  k.H = d->Hout;
  k.W = d->Wout;
  k.Cout = d->Cout;
  k.CinPad = round_up(d->Cin, 8);
  k.CoutPad = (int)(tcg_pack_geom(d->Cout, d->Cin, d->kd, d->kh, d->kw, 0).offset / ((long long)d->kd * d->kh * d->kw * k.CinPad));
  k.weight = d->weight;
  if (plan.khk) {
    const TcImg tk = tc_img_geom(d->Cout, d->Cin, d->kd, d->kh, d->kw, 0);
    ESM_REQUIRE(tk.kind == 1 && (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0, "conv(tc): no weight image for the kh-in-K kernel");
    k.wimg = d->weight + tk.offset;
    static const bool help_env = !(getenv("ESM_TC_TAILHELP") && atoi(getenv("ESM_TC_TAILHELP")) == 0);
    k.tail_help = help_env ? 1 : 0;
  } else if (!plan.gwc) {
    // the pack's image of this layer for tc_conv_kernel (same channel tile, taps and depth); the kernel stages the
    // weights itself when there is none (group-wise correlation stem: its 0.5 is folded into the weights)
    static const bool img_env = !(getenv("ESM_TC_WIMG") && atoi(getenv("ESM_TC_WIMG")) == 0);
    const TcImg tk = tc_img_geom(d->Cout, d->Cin, d->kd, d->kh, d->kw, 0);
    if ((img_env || plan.k1ts) && tk.kind == 2 && tk.COT == plan.COT && tk.taps == plan.taps && tk.KD == plan.KD &&
        (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0)
      k.wimg = d->weight + tk.offset;
    ESM_REQUIRE(!plan.k1ts || k.wimg, "conv(tc): no weight image for the pointwise TS-form kernel");
  }
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
  k.nseg = plan.nseg;
  k.segw = plan.segw;
  k.rows = plan.rows;
  k.nstrips = plan.nseg * plan.ysplit;
  k.groups = ceil_div(k.nstrips, 4);
  k.ztiles = ceil_div(d->Dout, plan.TZ);
  k.items_per_cot = d->B * k.ztiles * k.groups;
  k.ctas_per_cot = plan.ctas_per_cot;
  k.nstages = plan.nstages;
  k.npass = plan.npass;
  k.ps = d->pixel_shuffle;
  tc_fn_t fn = plan.k1ts ? tc1_pick(plan.COT) : plan.khk ? (plan.khk == 2 ? (tc_fn_t)tck_conv_kernel<32, true> : (tc_fn_t)tck_conv_kernel<32, false>) : tc_pick(plan.COT, plan.TZ, plan.KD, plan.gwc != 0, plan.taps, plan.NEW);
  ESM_REQUIRE(fn, "conv(tc): no kernel for COT=%d TZ=%d KD=%d", plan.COT, plan.TZ, plan.KD);
  // one limit for every launch of a function: the attribute is per function, not per launch, and graph
  // replays (and profilers re-launching graph nodes) must find it at least as large as any node's request
  if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return check_launch("conv(tc, cudaFuncSetAttribute)");
  launch_k(false, fn, dim3((unsigned)(plan.ncot * plan.ctas_per_cot)), dim3(tc_threads(plan.NEW)), plan.smem, st, k);
  ++tc_launches;
  return check_launch("conv(tc)");
}

}  // namespace esm

extern "C" long long esm_tc_conv_launches(void) { return esm::tc_launches; }
