// Instantiations of the direct-convolution kernel for kernel width 3, stride 1 (see conv_kernel.cuh).
#include "conv_kernel.cuh"

namespace esm {
conv_fn_t conv_kernels_k3(int COG, int CK, bool gwc, bool tma, int xo, int nv) {
  if (gwc) return nv == 4 ? (tma ? pick_gwc<true>(COG, CK) : pick_gwc<false>(COG, CK)) : nullptr;
  if (!tma) return pick_cog_ck<3, 1, false, 0>(COG, CK, nv);
  if (xo == 3) return pick_cog_ck<3, 1, true, 3>(COG, CK, nv);
  return nullptr;
}
}  // namespace esm
