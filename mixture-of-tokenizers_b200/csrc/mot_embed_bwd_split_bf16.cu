// Backward instantiations for the two halves of the split [tok | bytes] concat operand (MODE 2: tok-only rows with
// the token norm, MODE 3: bytes-only rows with the per-byte norm), element type __nv_bfloat16, rows of 256..1024.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_split_bf16(const EmbedParams& p, int mode, cudaStream_t s) {
  using T = __nv_bfloat16;
  const int cpl = p.Do / (32 * kBwdCW);  // exact: pick_mode checked Do % 128 == 0
  if (mode == 2) {
    switch (cpl) {
      case 2: return launch_bwd<T, 2, 2>(p, s);
      case 4: return launch_bwd<T, 4, 2>(p, s);
      case 6: return launch_bwd<T, 6, 2>(p, s);
      case 7: return launch_bwd<T, 7, 2>(p, s);
      case 8: return launch_bwd<T, 8, 2>(p, s);
    }
  } else {
    switch (cpl) {
      case 4: return launch_bwd<T, 4, 3>(p, s);
      case 6: return launch_bwd<T, 6, 3>(p, s);
      case 8: return launch_bwd<T, 8, 3>(p, s);
    }
  }
  return -1;  // no instantiation: the caller falls back to the run-time-flag kernel
}
}  // namespace mot
