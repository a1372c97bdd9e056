// Forward kernel instantiations, element type __nv_bfloat16.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_fwd_bf16(const EmbedParams& p, cudaStream_t s) {
  using T = __nv_bfloat16;
  switch ((p.n_chunks + 31) / 32) {
    case 1: return launch_fwd<T, 1>(p, s);
    case 2: return launch_fwd<T, 2>(p, s);
    case 3: return launch_fwd<T, 3>(p, s);
    case 4: return launch_fwd<T, 4>(p, s);
    case 5: case 6: return launch_fwd<T, 6>(p, s);
    case 7: case 8: return launch_fwd<T, 8>(p, s);
  }
  return MOT_ERR_UNSUPPORTED;
}
}  // namespace mot
