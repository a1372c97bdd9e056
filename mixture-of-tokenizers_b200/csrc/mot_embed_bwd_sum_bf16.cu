// Saved-output backward of the MoT-sum variant (mot_embed_bwd_sum.cuh), element type __nv_bfloat16.
#include "mot_embed_bwd_sum.cuh"
namespace mot {
int dispatch_bwd_sum_bf16(const EmbedParams& p, cudaStream_t s) {
  using T = __nv_bfloat16;
  if (!sum_path_ok(p)) return -1;
  switch (p.Do / (32 * kBwdCW)) {  // exact: pick_mode checked Do % 128 == 0
    case 4: return launch_bwd_sum<T, 4>(p, s);   // 512 = 8 x 64
    case 6: return launch_bwd_sum<T, 6>(p, s);   // 768 = 16 x 48
    case 8: return launch_bwd_sum<T, 8>(p, s);   // 1024 = 16 x 64 / 8 x 128 / 32 x 32
  }
  return -1;  // no instantiation: the caller runs the recompute kernel
}
}  // namespace mot
