// Backward instantiations of the ADD family with compile-time variant flags (MODE 16 + f, see Cfg), bf16:
// runs/73 (f = 3: per-input norms), runs/74 (f = 11: + lambdas), runs/71041..66 (f = 15: + output norm).
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_static_bf16(const EmbedParams& p, int mode, cudaStream_t s) {
  using T = __nv_bfloat16;
  const int cpl = p.Do / (32 * kBwdCW);  // exact: pick_mode checked Do % 128 == 0
  if (mode == 5) return dispatch_bwd_gather_bf16(p, s);
  if (mode == 6) return dispatch_bwd_concat_bf16(p, s);
#define MOT_STATIC_CASE(F)                                  \
  case 16 + F:                                              \
    if (cpl == 4) return launch_bwd<T, 4, 16 + F>(p, s);    \
    if (cpl == 8) return launch_bwd<T, 8, 16 + F>(p, s);    \
    break;
  switch (mode) {
    MOT_STATIC_CASE(3)
    MOT_STATIC_CASE(11)
    MOT_STATIC_CASE(15)
  }
#undef MOT_STATIC_CASE
  return -1;  // no instantiation: the caller runs the run-time-flag kernel
}
}  // namespace mot
