// Instantiations of the direct-convolution kernel for kernel width 5, stride 1 (see conv_kernel.cuh).
#include "conv_kernel.cuh"

namespace esm {
conv_fn_t conv_kernels_k5(int COG, int CK, bool gwc, bool tma, int xo, int nv) {
  if (gwc) return nullptr;
  if (nv != 4) return nullptr;  // k5 only ever sees single-channel inputs (CK=1)
  if (!tma) return pick_cog_ck<5, 1, false, 0>(COG, CK, nv);
  if (xo == 3) return pick_cog_ck<5, 1, true, 3>(COG, CK, nv);
  return nullptr;
}
}  // namespace esm
