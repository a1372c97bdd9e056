// Backward kernel instantiations (row widths up to 1024 elements), element type __nv_bfloat16.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_wide_bf16(const EmbedParams& p, cudaStream_t s);
int dispatch_bwd_bf16(const EmbedParams& p, cudaStream_t s) {
  using T = __nv_bfloat16;
  const int cpl = (p.n_chunks + 31) / 32;
  if (pick_mode(p) == 1) {  // MoT-sum fast path (runs/71)
    if (cpl == 3) return launch_bwd<T, 3, 1>(p, s);
    if (cpl == 4) return launch_bwd<T, 4, 1>(p, s);
  }
  switch (cpl) {
    case 1: return launch_bwd<T, 1, 0>(p, s);
    case 2: return launch_bwd<T, 2, 0>(p, s);
    case 3: return launch_bwd<T, 3, 0>(p, s);
    case 4: return launch_bwd<T, 4, 0>(p, s);
    default: return dispatch_bwd_wide_bf16(p, s);
  }
}
int launch_finalize_bf16(const EmbedParams& p, int blocks, cudaStream_t s) {
  mot_bwd_finalize_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(p);
  count_launch();
  return check_launch();
}
}  // namespace mot
