// Backward kernel instantiations (row widths 1025..2048 elements), element type __nv_bfloat16.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_wide_bf16(const EmbedParams& p, cudaStream_t s) {
  using T = __nv_bfloat16;
  const int cpl = (p.Do + 32 * kBwdCW - 1) / (32 * kBwdCW);
  if (cpl <= 12) return launch_bwd<T, 12, 0>(p, s);
  if (cpl <= 16) return launch_bwd<T, 16, 0>(p, s);
  return MOT_ERR_UNSUPPORTED;
}
}  // namespace mot
