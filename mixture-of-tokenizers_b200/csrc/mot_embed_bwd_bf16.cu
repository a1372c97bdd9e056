// Backward kernel instantiations (row widths up to 1024 elements), element type __nv_bfloat16.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_wide_bf16(const EmbedParams& p, cudaStream_t s);
int dispatch_bwd_bf16(const EmbedParams& p, cudaStream_t s) {
  using T = __nv_bfloat16;
  const int cpl = (p.Do + 32 * kBwdCW - 1) / (32 * kBwdCW);  // 4-element chunks per lane
  const int mode = pick_mode(p, kBwdCW);
  if (mode == 2 || mode == 3) {
    const int rc = dispatch_bwd_split_bf16(p, mode, s);
    if (rc >= 0) return rc;
  }
  if (mode == 5 || mode == 6 || mode >= 16) {  // plain gather / ADD family with compile-time flags
    const int rc = dispatch_bwd_static_bf16(p, mode, s);
    if (rc >= 0) return rc;
  }
  if (mode == 1) {  // MoT-sum fast path (runs/71): 512 = 8 x 64, 768 = 16 x 48, 1024 = 16 x 64 / 8 x 128 / 32 x 32
    if (cpl == 4) return launch_bwd<T, 4, 1>(p, s);
    if (cpl == 6) return launch_bwd<T, 6, 1>(p, s);
    if (cpl == 8) return launch_bwd<T, 8, 1>(p, s);
  }
  if (cpl <= 2) return launch_bwd<T, 2, 0>(p, s);
  if (cpl <= 4) return launch_bwd<T, 4, 0>(p, s);
  if (cpl <= 6) return launch_bwd<T, 6, 0>(p, s);
  if (cpl <= 8) return launch_bwd<T, 8, 0>(p, s);
  return dispatch_bwd_wide_bf16(p, s);
}
int launch_finalize_bf16(const EmbedParams& p, int blocks, cudaStream_t s) {
  launch_pdl(mot_bwd_finalize_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, s, p);
  count_launch();
  return check_launch();
}
}  // namespace mot
