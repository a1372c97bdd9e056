// Shared device/host helpers for libmot_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mot_b200.h"

namespace mot {

constexpr int kWarp = 32;
constexpr int kChunk = 8;  // elements handled by one lane per chunk (16 B of bf16 / 32 B of fp32)

// ---------------------------------------------------------------- host side
extern thread_local cudaError_t g_last_cuda_error;
void count_launch(int n = 1);
int check_launch();  // cudaGetLastError -> MOT_OK / MOT_ERR_CUDA
int device_props(int* sm_count, int* smem_optin);
// optional event pairs recorded around the main forward / backward kernel (mot_profile_events)
long long* chk_record();  // MOT_CHECK builds: host-mapped violation record (device pointer), else nullptr
extern long long* g_trace;  // device buffer for per-warp time stamps (mot_profile_trace; only libraries built with -DMOT_TRACE write it)
extern cudaEvent_t g_prof_fwd_start, g_prof_fwd_stop, g_prof_start, g_prof_stop;

// Launch with programmatic dependent launch (PDL): the kernel may be scheduled while its predecessor in the stream
// is still draining; every kernel of this library calls pdl_wait() before it touches global memory, so ordering and
// visibility are those of a normal stream launch while launch latency and CTA ramp-up overlap the predecessor's tail.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---------------------------------------------------------------- device side
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// per-warp time stamps for the timeline experiments (tools/trace_*.py): 64 slots of 8 bytes per warp
#ifdef MOT_TRACE
#define MOT_STAMP(buf, gw, slot)                                                                      \
  do {                                                                                               \
    if ((buf) != nullptr && (threadIdx.x & 31) == 0 && (slot) < 64) {                                 \
      unsigned long long t_;                                                                         \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      (buf)[(size_t)(gw) * 64 + (slot)] = (long long)t_;                                             \
    }                                                                                                \
  } while (0)
#else
#define MOT_STAMP(buf, gw, slot) do { } while (0)
#endif
// bounds checks of debug builds (-DMOT_CHECK): the first violation is written to a host-mapped record (it survives the
// dead context; a SIGABRT handler of the check build prints it) and the kernel traps.  Kernels reach the record through
// their parameter block (`p.chk`).
#ifdef MOT_CHECK
#define MOT_ASSERT(cond, what, a, b)                                                                              \
  do {                                                                                                           \
    if (!(cond) && p.chk != nullptr) {                                                                           \
      p.chk[1] = __LINE__; p.chk[2] = (long long)(a); p.chk[3] = (long long)(b);                                 \
      p.chk[4] = blockIdx.x; p.chk[5] = threadIdx.x; p.chk[6] = ((long long)p.v_lo << 32) | (unsigned)p.v_hi; p.chk[7] = ((long long)p.plan_early << 40) | ((long long)p.R << 20) | gridDim.x; p.chk[0] = 1;                                               \
      __threadfence_system();                                                                                    \
      __trap();                                                                                                  \
    }                                                                                                            \
  } while (0)
#else
#define MOT_ASSERT(cond, what, a, b) do { } while (0)
#endif
// PDL: let the next kernel in the stream start launching / block until every predecessor grid has completed and its
// memory is visible.  Both are no-ops when the kernel was launched without the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Global loads that STAY BELOW griddepcontrol.wait.  The __ldg() intrinsic is an invariant load to the compiler: it may
// hoist it above an `asm volatile(... ::: "memory")`, i.e. above pdl_wait(), and the kernel then reads what its
// predecessor in the programmatic-launch chain has not written yet (seen in round 2: off[V] of the sort plan read before
// plan_scan_kernel had stored it, once the launches were dense enough for the overlap to happen).  volatile asm
// statements keep their order among themselves, so these do not move across pdl_wait().
__device__ __forceinline__ int ld_g(const int* p) {
  int v;
  asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ long long ld_g(const long long* p) {
  long long v;
  asm volatile("ld.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_g(const float* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ short ld_g(const short* p) {
  short v;
  asm volatile("ld.global.s16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_g(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}

__device__ __forceinline__ uint4 ldg_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_16(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 lds_16(const void* p) { return *reinterpret_cast<const uint4*>(p); }

__device__ __forceinline__ void bf16x2_to_f32(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// CW consecutive elements <-> CW floats (CW = 8: 16 B of bf16 / 32 B of fp32; CW = 4: 8 B / 16 B).  Pointers must be
// aligned to the access size.  `Raw` is the register image of the elements as loaded (kept packed while a load is
// in flight).
__device__ __forceinline__ uint2 ldg_nc_8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_8(void* p, const uint2& v) {
  asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 lds_8(const void* p) { return *reinterpret_cast<const uint2*>(p); }

template <typename T, int CW>
struct Vec;

template <>
struct Vec<__nv_bfloat16, 8> {
  using Raw = uint4;
  __device__ __forceinline__ static Raw ldg_raw(const __nv_bfloat16* p) { return ldg_nc_16(p); }
  __device__ __forceinline__ static Raw lds_raw(const __nv_bfloat16* p) { return lds_16(p); }
  __device__ __forceinline__ static void unpack(const Raw& r, float (&v)[8]) {
    bf16x2_to_f32(r.x, v[0], v[1]);
    bf16x2_to_f32(r.y, v[2], v[3]);
    bf16x2_to_f32(r.z, v[4], v[5]);
    bf16x2_to_f32(r.w, v[6], v[7]);
  }
  __device__ __forceinline__ static void stg(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    r.x = f32x2_to_bf16x2(v[0], v[1]);
    r.y = f32x2_to_bf16x2(v[2], v[3]);
    r.z = f32x2_to_bf16x2(v[4], v[5]);
    r.w = f32x2_to_bf16x2(v[6], v[7]);
    stg_16(p, r);
  }
};

template <>
struct Vec<__nv_bfloat16, 4> {
  using Raw = uint2;
  __device__ __forceinline__ static Raw ldg_raw(const __nv_bfloat16* p) { return ldg_nc_8(p); }
  __device__ __forceinline__ static Raw lds_raw(const __nv_bfloat16* p) { return lds_8(p); }
  __device__ __forceinline__ static void unpack(const Raw& r, float (&v)[4]) {
    bf16x2_to_f32(r.x, v[0], v[1]);
    bf16x2_to_f32(r.y, v[2], v[3]);
  }
  __device__ __forceinline__ static void stg(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 r;
    r.x = f32x2_to_bf16x2(v[0], v[1]);
    r.y = f32x2_to_bf16x2(v[2], v[3]);
    stg_8(p, r);
  }
};

template <>
struct Vec<float, 8> {
  struct Raw {
    uint4 a, b;
  };
  __device__ __forceinline__ static Raw ldg_raw(const float* p) { return Raw{ldg_nc_16(p), ldg_nc_16(p + 4)}; }
  __device__ __forceinline__ static Raw lds_raw(const float* p) { return Raw{lds_16(p), lds_16(p + 4)}; }
  __device__ __forceinline__ static void unpack(const Raw& r, float (&v)[8]) {
    v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y); v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
    v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y); v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
  }
  __device__ __forceinline__ static void stg(float* p, const float (&v)[8]) {
    uint4 a, b;
    a.x = __float_as_uint(v[0]); a.y = __float_as_uint(v[1]); a.z = __float_as_uint(v[2]); a.w = __float_as_uint(v[3]);
    b.x = __float_as_uint(v[4]); b.y = __float_as_uint(v[5]); b.z = __float_as_uint(v[6]); b.w = __float_as_uint(v[7]);
    stg_16(p, a);
    stg_16(p + 4, b);
  }
};

template <>
struct Vec<float, 4> {
  using Raw = uint4;
  __device__ __forceinline__ static Raw ldg_raw(const float* p) { return ldg_nc_16(p); }
  __device__ __forceinline__ static Raw lds_raw(const float* p) { return lds_16(p); }
  __device__ __forceinline__ static void unpack(const Raw& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  __device__ __forceinline__ static void stg(float* p, const float (&v)[4]) {
    uint4 a;
    a.x = __float_as_uint(v[0]); a.y = __float_as_uint(v[1]); a.z = __float_as_uint(v[2]); a.w = __float_as_uint(v[3]);
    stg_16(p, a);
  }
};

template <typename T>
using Vec8 = Vec<T, 8>;

// ---- mbarrier + 1-D bulk (TMA) copy global -> shared, used to stage the byte table ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  // try_wait suspends the thread in hardware until the phase completes or a time limit passes; loop on the limit
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bytes must be a multiple of 16, both pointers 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace mot
