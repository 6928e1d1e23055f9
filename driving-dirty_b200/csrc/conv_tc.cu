// tcgen05 / TMEM implicit-GEMM convolutions (bf16) -- placeholder until the kernels land.
#include "dd_common.cuh"
namespace dd {
bool conv_tc_supported(int, int, int, int) { return false; }
int conv3x3_c32_fwd_tc(const void*, const float*, const float*, void*, int, int, int, int, int, const void*, cudaStream_t) {
  return fail(DD_ERR_UNSUPPORTED, "tcgen05 conv not built");
}
int conv3x3_c32_wgrad_tc(const void*, const void*, float*, float*, void*, size_t, int, int, int, int, cudaStream_t) {
  return fail(DD_ERR_UNSUPPORTED, "tcgen05 wgrad not built");
}
}  // namespace dd
