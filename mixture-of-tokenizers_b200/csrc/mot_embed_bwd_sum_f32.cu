// Saved-output backward of the MoT-sum variant (mot_embed_bwd_sum.cuh), element type float.
#include "mot_embed_bwd_sum.cuh"
namespace mot {
int dispatch_bwd_sum_f32(const EmbedParams& p, cudaStream_t s) {
  using T = float;
  if (!sum_path_ok(p)) return -1;
  switch (p.Do / (32 * kBwdCW)) {
    case 2: return launch_bwd_sum<T, 2>(p, s);
    case 4: return launch_bwd_sum<T, 4>(p, s);
    case 6: return launch_bwd_sum<T, 6>(p, s);
    case 8: return launch_bwd_sum<T, 8>(p, s);
  }
  return -1;
}
}  // namespace mot
