// Run-time-flag kernels with the dense addend (MODE 4: z = combine(...) + addend, d_addend = d z; runs/71051:226-229),
// element type float.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_fwd_addend_f32(const EmbedParams& p, cudaStream_t s) {
  using T = float;
  switch ((p.n_chunks + 31) / 32) {
    case 1: return launch_fwd<T, 1, 4>(p, s);
    case 2: return launch_fwd<T, 2, 4>(p, s);
    case 3: return launch_fwd<T, 3, 4>(p, s);
    case 4: return launch_fwd<T, 4, 4>(p, s);
  }
  return MOT_ERR_UNSUPPORTED;  // rows wider than 1024 elements
}
int dispatch_bwd_addend_f32(const EmbedParams& p, cudaStream_t s) {
  using T = float;
  const int cpl = (p.Do + 32 * kBwdCW - 1) / (32 * kBwdCW);
  if (cpl <= 2) return launch_bwd<T, 2, 4>(p, s);
  if (cpl <= 4) return launch_bwd<T, 4, 4>(p, s);
  if (cpl <= 6) return launch_bwd<T, 6, 4>(p, s);
  if (cpl <= 8) return launch_bwd<T, 8, 4>(p, s);
  return MOT_ERR_UNSUPPORTED;
}
}  // namespace mot
