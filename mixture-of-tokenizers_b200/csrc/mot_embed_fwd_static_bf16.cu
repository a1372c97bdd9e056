// Forward instantiations of the ADD family with compile-time variant flags (MODE 16 + f, see Cfg), bf16:
// runs/73 (f = 3), runs/74 (f = 11), runs/71041..66 (f = 15) at 512 and 1024 columns.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_fwd_static_bf16(const EmbedParams& p, int mode, cudaStream_t s) {
  using T = __nv_bfloat16;
  const int cpl = p.Do / (32 * 8);  // exact: pick_mode checked Do % 256 == 0
#define MOT_STATIC_CASE(F)                                  \
  case 16 + F:                                              \
    if (cpl == 2) return launch_fwd<T, 2, 16 + F>(p, s);    \
    if (cpl == 4) return launch_fwd<T, 4, 16 + F>(p, s);    \
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
