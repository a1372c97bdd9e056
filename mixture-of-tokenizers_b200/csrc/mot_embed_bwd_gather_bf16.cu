// Backward of a plain token gather (MODE 5: no norm, no lambdas -- the value embeddings, runs/7:252,308), bf16:
// d E[v] = sum of the upstream rows of v's occurrences; no token rows, no byte table.
#include "mot_embed_kernels.cuh"
namespace mot {
int dispatch_bwd_gather_bf16(const EmbedParams& p, cudaStream_t s) {
  using T = __nv_bfloat16;
  switch (p.Do / (32 * kBwdCW)) {
    case 2: return launch_bwd<T, 2, 5>(p, s);
    case 4: return launch_bwd<T, 4, 5>(p, s);
    case 6: return launch_bwd<T, 6, 5>(p, s);
    case 8: return launch_bwd<T, 8, 5>(p, s);
  }
  return -1;
}
// Pure concat under one output norm (MODE 6, runs/711:224-232: 512 + 16 x 32 = 1024 columns).
int dispatch_bwd_concat_bf16(const EmbedParams& p, cudaStream_t s) {
  using T = __nv_bfloat16;
  if (p.Do / (32 * kBwdCW) == 8) return launch_bwd<T, 8, 6>(p, s);
  return -1;
}
}  // namespace mot
