// Tensor-core (tcgen05 / TMEM) path of the convolution family: k3 s1 p1 convolutions, 2D and 3D,
// optionally with the group-wise correlation volume generated on the fly (see conv_tc.cu).
#pragma once
#include "common.cuh"

#include <cuda.h>  // CUtensorMap (types only)

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
  int rstages;       // raw (TMA-staged fp32) ring depth
  int ctas_per_cot;  // persistent CTAs per channel tile
  int npass;         // 3: split-TF32 (fp32-grade), 1: single-pass TF32
  size_t smem;
  int nstages_tma;   // operand ring depth when the inputs are TMA-staged
  size_t smem_tma;
};

// fp32 tensor map of rank `rank` (innermost first); strides in elements for dims 1..rank-1 (conv.cu)
bool encode_map(CUtensorMap* m, const void* base, int rank, const long long* dims, const long long* strides_elems, const int* box);

// Fills `plan` and returns true when `d` can run on the tensor-core path.
bool tc_conv_plan(const esm_conv_t* d, int num_sms, int npass, TcPlan* plan);
int tc_conv_launch(const esm_conv_t* d, const TcPlan& plan, cudaStream_t st);

}  // namespace esm
