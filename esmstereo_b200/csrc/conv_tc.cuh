// Tensor-core (tcgen05 / TMEM) path of the convolution family: k3 s1 p1 convolutions, 2D and 3D,
// optionally with the group-wise correlation volume generated on the fly (see conv_tc.cu).
#pragma once
#include "common.cuh"

namespace esm {

struct TcPlan {
  int COT;           // output channels per CTA (k3: 8, 16 or 24; k1: up to 64): N = taps*COT accumulator columns
  int taps;          // 9 (k3 s1 p1 in h,w) or 1 (pointwise)
  int TZ;            // output planes per work item (3 for COT=8 in 3D, else 1)
  int KD;            // 1 (2D) or 3
  int gwc;           // input voxels are group-wise correlations
  int ncot;          // output-channel tiles (= CTA groups; every CTA keeps one tile's weights resident)
  int nseg, segw;    // W is cut into nseg segments of segw (<= 30) output columns
  int ysplit, rows;  // H is cut into ysplit ranges of `rows` output rows
  int nstages;       // operand ring depth
  int ctas_per_cot;  // persistent CTAs per channel tile
  int npass;         // 3: split-TF32 (fp32-grade), 1: single-pass TF32
  int NEW;           // epilogue warps: 8 (two per TMEM lane quadrant) or 16
  int khk;           // 2D k3, Cout % 32 == 0: kh in K, kw in N, 32 channels per CTA (tck_conv_kernel; 2 = A in tensor memory)
  int k1ts;          // pointwise: TS-form kernel over flat pixel tiles (tc1_conv_kernel)
  size_t smem;
};

// Geometry of the split-TF32 weight slabs stored after the fp32 pack (conv.cu: esm_pack_conv_weight_f32).
struct TcgPack {
  long long offset, elems;  // in floats, from the start of the packed weight
  int phases, taps, KD, KH, KW, ncg, CoutX;
};
TcgPack tcg_pack_geom(int Cout, int Cin, int kd, int kh, int kw, int transposed);

// Third region of a packed weight (stride-1 k1 / k3 layers with Cin >= 8): the resident-weight image of the tcgen05
// kernels of conv_tc.cu, split into hi / lo and laid out exactly as the kernel keeps it in shared memory, so that a CTA's
// weights are a few bulk copies instead of thousands of clocks of scalar loads, index arithmetic and stores:
//   kind 1 (2D k3, Cout % 32 == 0; tck_conv_kernel):  [Cout / 32][cg][kh][hi | lo][k / 4][n = (co % 32) * 3 + kw][k % 4]
//   kind 2 (every other k1 / k3 layer; tc_conv_kernel): [cot][cg][kd][hi | lo][k / 4][n = (co % COT) * taps + tap][k % 4]
// with COT the channel tile tc_conv_plan uses for that Cout (tc_cot()).
struct TcImg {
  int kind;                 // 0: none
  long long offset, elems;  // in floats, from the start of the packed weight (offset is a multiple of 32)
  int COT, taps, KD, ncot, ncg;
  long long per_cot;        // floats per channel tile
};
int tc_cot(int Cout, bool k1);  // output channels per CTA of tc_conv_kernel, 0 = not eligible
TcImg tc_img_geom(int Cout, int Cin, int kd, int kh, int kw, int transposed);

// Streamed-weight GEMM path (conv_tcg.cu): any k / stride 1-2 / transposed k4 s2 layer with Cin >= 8.
struct TcgPlan {
  int NT;        // output channels (accumulator columns) per CTA, multiple of 8, <= 128
  int ncot;      // channel tiles
  int mtiles;    // 128-voxel tiles of the output lattice (per batch item and phase)
  int nstages;   // operand ring depth
  int ctas;      // persistent CTAs
  int npass;
  size_t smem;
};
bool tcg_conv_plan(const esm_conv_t* d, int num_sms, int npass, TcgPlan* plan);
int tcg_conv_launch(const esm_conv_t* d, const TcgPlan& plan, cudaStream_t st);

// Streaming kernel for pointwise (k1) layers (conv_pw.cu): true fp32, HBM-bound.
struct PwPlan {
  int CO;       // output channels per thread (8 / 16 / 24 / 32)
  int cotiles;  // channel tiles (gridDim.y)
  size_t smem;
};
bool pw_conv_plan(const esm_conv_t* d, PwPlan* plan);
int pw_conv_launch(const esm_conv_t* d, const PwPlan& plan, cudaStream_t st);

// Dedicated kernel for the 3 -> C k3 stride-2 image-side layers (conv_stem3.cu): exact fp32, taken whenever eligible.
bool stem3_eligible(const esm_conv_t* d);
int stem3_launch(const esm_conv_t* d, cudaStream_t st);

// Fills `plan` and returns true when `d` can run on the tensor-core path.
bool tc_conv_plan(const esm_conv_t* d, int num_sms, int npass, TcPlan* plan);
int tc_conv_launch(const esm_conv_t* d, const TcPlan& plan, cudaStream_t st);

}  // namespace esm
