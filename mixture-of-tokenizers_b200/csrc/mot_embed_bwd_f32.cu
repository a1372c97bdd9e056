// Backward kernel instantiations (row widths up to 1024 elements), element type float.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_wide_f32(const EmbedParams& p, cudaStream_t s);
int dispatch_bwd_f32(const EmbedParams& p, cudaStream_t s) {
  using T = float;
  const int cpl = (p.Do + 32 * kBwdCW - 1) / (32 * kBwdCW);
  if (cpl <= 2) return launch_bwd<T, 2, 0>(p, s);
  if (cpl <= 4) return launch_bwd<T, 4, 0>(p, s);
  if (cpl <= 6) return launch_bwd<T, 6, 0>(p, s);
  if (cpl <= 8) return launch_bwd<T, 8, 0>(p, s);
  return dispatch_bwd_wide_f32(p, s);
}
int launch_finalize_f32(const EmbedParams& p, int blocks, cudaStream_t s) {
  launch_pdl(mot_bwd_finalize_kernel<float>, dim3(blocks), dim3(256), 0, s, p);
  count_launch();
  return check_launch();
}
}  // namespace mot
