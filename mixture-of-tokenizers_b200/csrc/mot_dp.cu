// The one exchange step of the path: average the dense gradients of the replicated tables across the data-parallel
// ranks (reference: one dist.all_reduce(param.grad, AVG) per parameter, spt/train_gpt.py:1320-1321, runs/7:697-700).
//
// The flat gradient bucket the backward kernels wrote lives in symmetric memory that is also mapped as ONE multicast
// address range over all ranks' copies (NVLink 5 / NVSwitch, NVLS).  One kernel per rank, two-shot through the switch:
//   barrier (signal pads, release/acquire at system scope): every rank's backward has finished writing its copy
//   rank r owns slice r of the bucket: multimem.ld_reduce pulls that slice from ALL copies, summed inside the switch
//     with fp32 accumulation; scale by 1/world; multimem.st pushes the averaged slice back to ALL copies
//   barrier: every slice has landed everywhere before any rank reads the bucket
// Per GPU and direction about one bucket size crosses NVLink, independent of the number of ranks.
#include <cstdlib>

#include "mot_common.cuh"

namespace mot {

// Few, fat CTAs: the switch round trip is microseconds and the links saturate with about 1 MB in flight per rank; more
// CTAs only add contention inside the switch (8 ranks, 77 MB bf16: 8 x 1024 threads 195 us, 16 x 1024 203 us, 36 x 1024
// 219 us, 144 x 512 239 us, NCCL 270 us; profiles/r1_experiments.md).
constexpr int kArThreads = 1024;
constexpr int kArMaxBlocks = 8;     // signal-pad slots used: blocks x world uint32 (torch's pad is 9216 B = 2304 slots)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Block b of every rank meets block b of every other rank: lane r tells rank r "rank `rank` reached `target`", then
// waits until rank r said the same.  Counters only grow, so nothing is ever reset.
__device__ __forceinline__ void rank_barrier(uint32_t* const* pads, int rank, int world, uint32_t target) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pads[threadIdx.x] + blockIdx.x * world + rank, target);
    const uint32_t* mine = pads[rank] + blockIdx.x * world + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - target) < 0) {
    }
  }
  __syncthreads();
}

template <bool BF16, int kUnroll>
__global__ void __launch_bounds__(1024) nvls_allreduce_avg_kernel(char* mc, uint32_t* const* pads, int rank, int world,
                                                                        long long n_vec /* 16-byte vectors */, uint32_t epoch) {
  pdl_launch_dependents();
  pdl_wait();  // the local backward / finalize kernels have completed: this rank's copy is final
  rank_barrier(pads, rank, world, 2u * epoch + 1u);
  const long long per = (n_vec + world - 1) / world;
  const long long lo = per * rank, hi = min(lo + per, n_vec);
  const float inv = 1.f / (float)world;
  // kUnroll independent 16-byte reductions in flight per thread: one switch round trip is microseconds, the link wants
  // megabytes outstanding
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * kUnroll) {
    uint32_t r[kUnroll][4];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi) {
        char* a = mc + i * 16;
        if (BF16)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                       : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "l"(a) : "memory");
        else
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "l"(a) : "memory");
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (BF16) {
            float lo_f, hi_f;
            bf16x2_to_f32(r[u][j], lo_f, hi_f);
            r[u][j] = f32x2_to_bf16x2(lo_f * inv, hi_f * inv);
          } else {
            r[u][j] = __float_as_uint(__uint_as_float(r[u][j]) * inv);
          }
        }
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + i * 16), "r"(r[u][0]), "r"(r[u][1]),
                     "r"(r[u][2]), "r"(r[u][3])
                     : "memory");
      }
    }
  }
  rank_barrier(pads, rank, world, 2u * epoch + 2u);
}

}  // namespace mot

using namespace mot;

extern "C" int mot_dp_allreduce_avg(void* multicast_ptr, void* const* signal_pads_dev, int32_t rank, int32_t world, int64_t n_bytes,
                                    int32_t dtype, uint32_t epoch, void* stream) {
  if (!multicast_ptr || !signal_pads_dev || world < 1 || rank < 0 || rank >= world || n_bytes < 0) return MOT_ERR_BAD_ARG;
  if (dtype != MOT_BF16 && dtype != MOT_F32) return MOT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(multicast_ptr) & 15u) || (n_bytes & 15)) return MOT_ERR_MISALIGNED;
  if (world > 16) return MOT_ERR_UNSUPPORTED;
  if (n_bytes == 0) return MOT_OK;
  int sms = 0, optin = 0;
  if (int rc = device_props(&sms, &optin)) return rc;
  const long long n_vec = n_bytes / 16;
  const char* env_b = getenv("MOT_AR_BLOCKS");    // debug knobs, read per call so that one process can sweep them
  const char* env_t = getenv("MOT_AR_THREADS");
  const char* env_u = getenv("MOT_AR_UNROLL");
  int threads = env_t ? atoi(env_t) : kArThreads;
  if (threads < 32 || threads > 1024 || threads % 32) threads = kArThreads;
  const int unroll = env_u ? atoi(env_u) : 8;
  long long max_blocks = env_b ? atoi(env_b) : kArMaxBlocks;
  if (max_blocks < 1) max_blocks = 1;
  if (max_blocks * world * 4 > 9216) max_blocks = 9216 / (world * 4);   // signal-pad slots
  long long blocks = (n_vec / world + threads * 8 - 1) / (threads * 8);
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  uint32_t* const* pads = reinterpret_cast<uint32_t* const*>(signal_pads_dev);
  char* mc = reinterpret_cast<char*>(multicast_ptr);
  const dim3 g((unsigned)blocks), b(threads);
  if (dtype == MOT_BF16) {
    if (unroll == 4) launch_pdl(nvls_allreduce_avg_kernel<true, 4>, g, b, 0, s, mc, pads, (int)rank, (int)world, n_vec, epoch);
    else launch_pdl(nvls_allreduce_avg_kernel<true, 8>, g, b, 0, s, mc, pads, (int)rank, (int)world, n_vec, epoch);
  } else {
    if (unroll == 4) launch_pdl(nvls_allreduce_avg_kernel<false, 4>, g, b, 0, s, mc, pads, (int)rank, (int)world, n_vec, epoch);
    else launch_pdl(nvls_allreduce_avg_kernel<false, 8>, g, b, 0, s, mc, pads, (int)rank, (int)world, n_vec, epoch);
  }
  count_launch();
  return check_launch();
}
