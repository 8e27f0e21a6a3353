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
  size_t smem;
};

// Geometry of the split-TF32 weight slabs stored after the fp32 pack (conv.cu: esm_pack_conv_weight_f32).
struct TcgPack {
  long long offset, elems;  // in floats, from the start of the packed weight
  int phases, taps, KD, KH, KW, ncg, CoutX;
};
TcgPack tcg_pack_geom(int Cout, int Cin, int kd, int kh, int kw, int transposed);

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
