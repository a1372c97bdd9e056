// Backward kernel instantiations (row widths 1025..2048 elements), element type float.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_wide_f32(const EmbedParams& p, cudaStream_t s) {
  using T = float;
  switch ((p.n_chunks + 31) / 32) {
    case 5: case 6: return launch_bwd<T, 6, 0>(p, s);
    case 7: case 8: return launch_bwd<T, 8, 0>(p, s);
  }
  return MOT_ERR_UNSUPPORTED;
}
}  // namespace mot
