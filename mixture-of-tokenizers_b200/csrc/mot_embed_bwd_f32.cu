// Backward kernel instantiations (row widths up to 1024 elements), element type float.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_wide_f32(const EmbedParams& p, cudaStream_t s);
int dispatch_bwd_f32(const EmbedParams& p, cudaStream_t s) {
  using T = float;
  const int cpl = (p.n_chunks + 31) / 32;
  switch (cpl) {
    case 1: return launch_bwd<T, 1, 0>(p, s);
    case 2: return launch_bwd<T, 2, 0>(p, s);
    case 3: return launch_bwd<T, 3, 0>(p, s);
    case 4: return launch_bwd<T, 4, 0>(p, s);
    default: return dispatch_bwd_wide_f32(p, s);
  }
}
int launch_finalize_f32(const EmbedParams& p, int blocks, cudaStream_t s) {
  mot_bwd_finalize_kernel<float><<<blocks, 256, 0, s>>>(p);
  count_launch();
  return check_launch();
}
}  // namespace mot
