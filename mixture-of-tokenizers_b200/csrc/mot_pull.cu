// Byte "pull" across tokens, the step between the ttb expansion and the embedding gather.
//
// Replaces pull_from_left / pull_from_right (spt/data_creation.py:179-305 / :71-176; runs/7:351-428), which the
// reference runs as torch ops with a Python loop over batch rows, nonzero / searchsorted and host syncs.
//   pull_from_left  (rows left-padded):  a non-EOT token t receives the LAST min(bpt, n) of the n non-pad bytes of all
//                    tokens after the previous EOT (or row start) up to and including t, right-aligned, pad 456 left.
//   pull_from_right (rows right-padded): token t receives the FIRST min(bpt, n) of the non-pad bytes of tokens
//                    t, t+1, ... before the next EOT / row end, left-aligned.
//   A token is EOT iff all its bpt bytes equal eot_byte; EOT tokens pass through unchanged and reset the pool.
// pull_from_right is pull_from_left on the sequence read backwards (tokens reversed inside the row, bytes reversed
// inside the token), so one set of kernels serves both through an index map.
//
// Four small launches, no host reads: (1) per token: non-pad count and EOT / row-start flag, per 256-token block the
// byte total and the last flagged token; (2) one CTA scans the blocks; (3) per token: position in the compacted byte
// stream, compacted bytes written; (4) per token: segment start by a prefix max, output row.  Integer work, bit-exact.
#include "mot_common.cuh"

namespace mot {

constexpr int kPullThreads = 256;

struct PullParams {
  const void* in;
  void* out;
  int* cntflag;     // [n_tok]  non-pad count | flag << 8 | eot << 9   (logical order)
  int* off;         // [n_tok]  exclusive prefix sum of the counts = position in the compacted stream
  short* comp;      // [n_tok * bpt] compacted non-pad bytes (logical order)
  int* bsum;        // [n_blk]  bytes per block -> exclusive offsets after the block scan
  int* blast;       // [n_blk]  last flagged logical token of the block, -1 -> last flagged token before the block
  long long n_tok, T;
  int bpt, pad, eot, i64, rev, n_blk;
};

template <typename IdT>
__device__ __forceinline__ int ld_id(const PullParams& p, long long tok_phys, int slot_phys) {
  return (int)ld_g(reinterpret_cast<const IdT*>(p.in) + tok_phys * p.bpt + slot_phys);
}
// logical token -> physical token (reversed inside its row for pull_from_right)
__device__ __forceinline__ long long phys_tok(const PullParams& p, long long L) {
  if (!p.rev) return L;
  const long long row = L / p.T, t = L - row * p.T;
  return row * p.T + (p.T - 1 - t);
}
__device__ __forceinline__ int phys_slot(const PullParams& p, int j) { return p.rev ? p.bpt - 1 - j : j; }

__device__ __forceinline__ int block_excl_scan(int v, int* smem, int& total) {  // 256 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  int wo = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kPullThreads / 32; ++w) {
    const int s = smem[w];
    if (w < warp) wo += s;
    tot += s;
  }
  __syncthreads();
  total = tot;
  return wo + inc - v;
}
__device__ __forceinline__ int block_incl_max(int v, int* smem) {  // 256 threads, inclusive prefix max
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int m = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, m, o);
    if (lane >= o) m = max(m, t);
  }
  if (lane == 31) smem[warp] = m;
  __syncthreads();
  int wm = -1;
#pragma unroll
  for (int w = 0; w < kPullThreads / 32; ++w)
    if (w < warp) wm = max(wm, smem[w]);
  __syncthreads();
  return max(m, wm);
}

template <typename IdT>
__global__ void __launch_bounds__(kPullThreads) pull_count_kernel(PullParams p) {
  __shared__ int sm[kPullThreads / 32];
  pdl_launch_dependents();
  pdl_wait();
  const long long L = (long long)blockIdx.x * kPullThreads + threadIdx.x;
  int cnt = 0, flag = 0, eot = 0;
  if (L < p.n_tok) {
    const long long pt = phys_tok(p, L);
    int n_eot = 0;
    for (int j = 0; j < p.bpt; ++j) {
      const int id = ld_id<IdT>(p, pt, j);
      cnt += id != p.pad;
      n_eot += id == p.eot;
    }
    eot = n_eot == p.bpt;
    if (eot) cnt = 0;  // EOT tokens contribute nothing to the pool
    flag = eot || (L % p.T) == 0;
    p.cntflag[L] = cnt | (flag << 8) | (eot << 9);
  }
  int total;
  block_excl_scan(cnt, sm, total);
  const int last = block_incl_max(flag ? (int)(L - (long long)blockIdx.x * kPullThreads) : -1, sm);
  if (threadIdx.x == kPullThreads - 1) {
    p.bsum[blockIdx.x] = total;
    p.blast[blockIdx.x] = last;  // block-local index or -1
  }
}

// one CTA: exclusive sum of the block totals, and for every block the last flagged token of any earlier block
__global__ void __launch_bounds__(1024) pull_scan_blocks_kernel(PullParams p) {
  __shared__ int wsum[32], wmax[32];
  __shared__ int carry_sum, carry_max;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    carry_sum = 0;
    carry_max = -1;
  }
  __syncthreads();
  for (int base = 0; base < p.n_blk; base += 1024) {
    const int b = base + threadIdx.x;
    const int v = b < p.n_blk ? p.bsum[b] : 0;
    const int l = b < p.n_blk ? p.blast[b] : -1;
    const int g = l >= 0 ? b * kPullThreads + l : -1;  // global logical index (n_tok < 2^31)
    int inc = v, m = g;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      const int u = __shfl_up_sync(0xffffffffu, m, o);
      if (lane >= o) {
        inc += t;
        m = max(m, u);
      }
    }
    if (lane == 31) {
      wsum[warp] = inc;
      wmax[warp] = m;
    }
    __syncthreads();
    int wo = carry_sum, wm = carry_max, tot = 0, totm = -1;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) {
        wo += wsum[w];
        wm = max(wm, wmax[w]);
      }
      tot += wsum[w];
      totm = max(totm, wmax[w]);
    }
    // exclusive results: everything strictly before block b
    const int m_prev = __shfl_up_sync(0xffffffffu, m, 1);
    if (b < p.n_blk) {
      p.bsum[b] = wo + inc - v;
      p.blast[b] = max(wm, lane > 0 ? m_prev : -1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      carry_sum += tot;
      carry_max = max(carry_max, totm);
    }
    __syncthreads();
  }
}

template <typename IdT>
__global__ void __launch_bounds__(kPullThreads) pull_compact_kernel(PullParams p) {
  __shared__ int sm[kPullThreads / 32];
  pdl_launch_dependents();
  pdl_wait();
  const long long L = (long long)blockIdx.x * kPullThreads + threadIdx.x;
  const int cf = L < p.n_tok ? p.cntflag[L] : 0;
  const int cnt = cf & 0xff;
  int total;
  const int off = p.bsum[blockIdx.x] + block_excl_scan(cnt, sm, total);
  if (L < p.n_tok) {
    p.off[L] = off;
    if (cnt > 0) {
      const long long pt = phys_tok(p, L);
      int w = off;
      for (int j = 0; j < p.bpt; ++j) {
        const int id = ld_id<IdT>(p, pt, phys_slot(p, j));
        if (id != p.pad) p.comp[w++] = (short)id;
      }
    }
  }
}

template <typename IdT>
__global__ void __launch_bounds__(kPullThreads) pull_emit_kernel(PullParams p) {
  __shared__ int sm[kPullThreads / 32];
  pdl_launch_dependents();
  pdl_wait();
  const long long L = (long long)blockIdx.x * kPullThreads + threadIdx.x;
  const bool in = L < p.n_tok;
  const int cf = in ? p.cntflag[L] : 0;
  const int cnt = cf & 0xff, flag = (cf >> 8) & 1, eot = (cf >> 9) & 1;
  const int off = in ? p.off[L] : 0;
  // start of the current segment in the compacted stream: off of the last flagged token at or before L
  const int prev = p.blast[blockIdx.x];
  const int carry = prev >= 0 ? p.off[prev] : 0;
  const int seg = max(block_incl_max(flag ? off : -1, sm), carry);
  if (!in) return;
  const long long pt = phys_tok(p, L);
  IdT* dst = reinterpret_cast<IdT*>(p.out) + pt * p.bpt;
  if (eot) {
    for (int j = 0; j < p.bpt; ++j) dst[j] = (IdT)p.eot;
    return;
  }
  const int end = off + cnt;
  const int n = min(p.bpt, end - seg);
  for (int j = 0; j < p.bpt; ++j) {  // logical slot j: pads first, then the last n bytes of the pool
    const int id = j < p.bpt - n ? p.pad : (int)p.comp[end - p.bpt + j];
    dst[phys_slot(p, j)] = (IdT)id;
  }
}

struct PullWs {
  size_t cntflag, off, comp, bsum, blast, total;
};
static PullWs pull_ws(long long n_tok, int bpt) {
  PullWs w{};
  const long long n_blk = (n_tok + kPullThreads - 1) / kPullThreads;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = (o + bytes + 255) / 256 * 256;
    return at;
  };
  w.cntflag = take((size_t)n_tok * 4);
  w.off = take((size_t)n_tok * 4);
  w.comp = take((size_t)n_tok * bpt * 2);
  w.bsum = take((size_t)n_blk * 4);
  w.blast = take((size_t)n_blk * 4);
  w.total = o;
  return w;
}

template <typename IdT>
static int run_pull(PullParams& p, cudaStream_t s) {
  launch_pdl(pull_count_kernel<IdT>, dim3(p.n_blk), dim3(kPullThreads), 0, s, p);
  launch_pdl(pull_scan_blocks_kernel, dim3(1), dim3(1024), 0, s, p);
  launch_pdl(pull_compact_kernel<IdT>, dim3(p.n_blk), dim3(kPullThreads), 0, s, p);
  launch_pdl(pull_emit_kernel<IdT>, dim3(p.n_blk), dim3(kPullThreads), 0, s, p);
  count_launch(4);
  return check_launch();
}

}  // namespace mot

using namespace mot;

extern "C" size_t mot_pull_workspace_bytes(int64_t n_rows, int64_t tokens_per_row, int32_t bpt) {
  if (n_rows <= 0 || tokens_per_row <= 0 || bpt <= 0 || bpt > 255) return 0;
  return pull_ws(n_rows * tokens_per_row, bpt).total;
}

extern "C" int mot_pull(const void* bytes_in, void* bytes_out, int64_t n_rows, int64_t tokens_per_row, int32_t bpt,
                        int32_t ids_i64, int32_t pad_byte, int32_t eot_byte, int32_t from_right, void* workspace,
                        size_t ws_bytes, void* stream) {
  if (n_rows < 0 || tokens_per_row < 0 || bpt <= 0 || bpt > 255) return MOT_ERR_BAD_ARG;
  const long long n_tok = n_rows * tokens_per_row;
  if (n_tok == 0) return MOT_OK;
  if (n_tok > 0x7fffffffLL / bpt) return MOT_ERR_BAD_ARG;  // positions in the compacted stream are int32
  if (!bytes_in || !bytes_out || !workspace) return MOT_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) return MOT_ERR_MISALIGNED;
  const PullWs w = pull_ws(n_tok, bpt);
  if (ws_bytes < w.total) return MOT_ERR_WORKSPACE;
  int sms = 0, optin = 0;
  if (int rc = device_props(&sms, &optin)) return rc;
  char* b = reinterpret_cast<char*>(workspace);
  PullParams p{};
  p.in = bytes_in; p.out = bytes_out;
  p.cntflag = reinterpret_cast<int*>(b + w.cntflag);
  p.off = reinterpret_cast<int*>(b + w.off);
  p.comp = reinterpret_cast<short*>(b + w.comp);
  p.bsum = reinterpret_cast<int*>(b + w.bsum);
  p.blast = reinterpret_cast<int*>(b + w.blast);
  p.n_tok = n_tok; p.T = tokens_per_row; p.bpt = bpt; p.pad = pad_byte; p.eot = eot_byte; p.i64 = ids_i64; p.rev = from_right ? 1 : 0;
  p.n_blk = (int)((n_tok + kPullThreads - 1) / kPullThreads);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return ids_i64 ? run_pull<long long>(p, s) : run_pull<int>(p, s);
}
