// Forward kernel instantiations, element type __nv_bfloat16.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_fwd_bf16(const EmbedParams& p, cudaStream_t s) {
  using T = __nv_bfloat16;
  const int cpl = (p.n_chunks + 31) / 32;
  const int mode = pick_mode(p, 8);
  if (mode >= 16) {  // ADD family with compile-time flags
    const int rc = dispatch_fwd_static_bf16(p, mode, s);
    if (rc >= 0) return rc;
  }
  if (mode == 1) {  // MoT-sum fast path (runs/71), the shapes the reference ships
    if (cpl == 3) return launch_fwd<T, 3, 1>(p, s);
    if (cpl == 4) return launch_fwd<T, 4, 1>(p, s);
  }
  switch (cpl) {
    case 1: return launch_fwd<T, 1, 0>(p, s);
    case 2: return launch_fwd<T, 2, 0>(p, s);
    case 3: return launch_fwd<T, 3, 0>(p, s);
    case 4: return launch_fwd<T, 4, 0>(p, s);
    case 5: case 6: return launch_fwd<T, 6, 0>(p, s);
    case 7: case 8: return launch_fwd<T, 8, 0>(p, s);
  }
  return MOT_ERR_UNSUPPORTED;
}
}  // namespace mot
