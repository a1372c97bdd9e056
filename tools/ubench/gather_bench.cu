// Micro-benchmark: random-row gather read throughput on B200 with three fetch mechanisms
// (design evidence for the row fetch of mot_fwd_kernel / mot_bwd_kernel, not product code).
//   A: 1-D bulk async copy (cp.async.bulk, TMA engine), one op per row, per-warp ring + mbarriers
//   B: LDG.128 straight to registers, U rows in flight per warp
//   C: cp.async 16 B per lane (LDGSTS) into a per-warp ring, completion through mbarriers
// Rows are consumed (xor) so nothing is optimised away; no output stream.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ---- A: bulk copies.  split = number of bulk ops per row (1, 2, 4)
template <int STAGES>
__global__ void __launch_bounds__(1024, 1) k_bulk(const __nv_bfloat16* tab, const int* idx, int n, int D, int split, uint32_t* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t row_bytes = D * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * STAGES;
  unsigned char* ring = smem + 1024 * 8 + (size_t)warp * STAGES * row_bytes;
  if (lane == 0) for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int gw = warp * gridDim.x + blockIdx.x, W = gridDim.x * nw;
  const int n_i = gw < n ? (n - gw + W - 1) / W : 0;
  // indices: lane j holds the index of iteration 32*b + j for batch b (current and next batch), one coalesced load each
  auto load_batch = [&](int b) { const int i = b * 32 + lane; return i < n_i ? __ldg(idx + gw + i * W) : 0; };
  int cur = load_batch(0), nxt = load_batch(1);
  int cb = 0;  // batch held in `cur`
  auto issue = [&](int i, int s) {  // warp collective; i is in batch i/32 == (current or next)
    const int r = __shfl_sync(0xffffffffu, (i >> 5) == cb ? cur : nxt, i & 31);
    if (lane == 0) {
      mbar_expect_tx(bars + s, row_bytes);
      const uint32_t piece = row_bytes / split;
      for (int q = 0; q < split; ++q)
        bulk_g2s(ring + (size_t)s * row_bytes + q * piece, reinterpret_cast<const char*>(tab) + (size_t)r * row_bytes + q * piece, piece, bars + s);
    }
  };
  for (int i = 0; i < STAGES && i < n_i; ++i) issue(i, i);
  uint32_t acc = 0, par = 0;
  int s = 0;
  for (int i = 0; i < n_i; ++i) {
    mbar_wait(bars + s, par);
    const uint4* row = reinterpret_cast<const uint4*>(ring + (size_t)s * row_bytes);
    for (int c = lane; c < (int)row_bytes / 16; c += 32) { uint4 v = row[c]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    __syncwarp();
    if (i + STAGES < n_i) issue(i + STAGES, s);
    if (++s == STAGES) { s = 0; par ^= 1; }
    if (((i + STAGES) & 31) == 31) {  // the issue cursor leaves batch (i+STAGES)/32: rotate
      cur = nxt; ++cb; nxt = load_batch(cb + 1);
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- B: LDG.128 to registers, U rows in flight, CPL 16-byte chunks per lane per row (indices batch-prefetched)
template <int U, int CPL>
__global__ void __launch_bounds__(1024, 1) k_ldg(const __nv_bfloat16* tab, const int* idx, int n, int D, uint32_t* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gw = warp * gridDim.x + blockIdx.x, W = gridDim.x * nw;
  const int n_i = gw < n ? (n - gw + W - 1) / W : 0;
  auto load_batch = [&](int b) { const int i = b * 32 + lane; return i < n_i ? __ldg(idx + gw + i * W) : 0; };
  int cur = load_batch(0), nxt = load_batch(1);
  uint32_t acc = 0;
  for (int i = 0; i < n_i; i += U) {   // 32 % U == 0: a group of U rows never straddles a batch
    uint4 v[U][CPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = __shfl_sync(0xffffffffu, cur, (i + u) & 31);
      const char* row = reinterpret_cast<const char*>(tab) + (size_t)r * D * 2;
#pragma unroll
      for (int c = 0; c < CPL; ++c) v[u][c] = ldg_nc(row + (c * 32 + lane) * 16);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc ^= v[u][c].x ^ v[u][c].y ^ v[u][c].z ^ v[u][c].w;
    if (((i + U) & 31) == 0) { cur = nxt; nxt = load_batch(((i + U) >> 5) + 1); }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- C: cp.async 16 B per lane into a per-warp ring; completion via mbarrier (cp.async.mbarrier.arrive.noinc)
template <int STAGES, int CPL>
__global__ void __launch_bounds__(1024, 1) k_cpasync(const __nv_bfloat16* tab, const int* idx, int n, int D, uint32_t* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t row_bytes = D * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * STAGES;
  unsigned char* ring = smem + 1024 * 8 + (size_t)warp * STAGES * row_bytes;
  if (lane == 0) for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 32);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int gw = warp * gridDim.x + blockIdx.x, W = gridDim.x * nw;
  const int n_i = gw < n ? (n - gw + W - 1) / W : 0;
  auto load_batch = [&](int b) { const int i = b * 32 + lane; return i < n_i ? __ldg(idx + gw + i * W) : 0; };
  int cur = load_batch(0), nxt = load_batch(1);
  int cb = 0;
  auto issue = [&](int i, int s) {
    const int r = __shfl_sync(0xffffffffu, (i >> 5) == cb ? cur : nxt, i & 31);
    const char* row = reinterpret_cast<const char*>(tab) + (size_t)r * row_bytes;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int o = (c * 32 + lane) * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ring + (size_t)s * row_bytes + o)), "l"(row + o) : "memory");
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bars + s)) : "memory");
  };
  for (int i = 0; i < STAGES && i < n_i; ++i) issue(i, i);
  uint32_t acc = 0, par = 0;
  int s = 0;
  for (int i = 0; i < n_i; ++i) {
    mbar_wait(bars + s, par);
    const uint4* row = reinterpret_cast<const uint4*>(ring + (size_t)s * row_bytes);
#pragma unroll
    for (int c = 0; c < CPL; ++c) { uint4 v = row[c * 32 + lane]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    __syncwarp();
    if (i + STAGES < n_i) issue(i + STAGES, s);
    if (++s == STAGES) { s = 0; par ^= 1; }
    if (((i + STAGES) & 31) == 31) { cur = nxt; ++cb; nxt = load_batch(cb + 1); }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <typename F>
static float time_us(F f, int reps = 10) {
  for (int i = 0; i < 3; ++i) f();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  return ms / reps * 1e3f;
}

template <int CPL>
static void run(int D, int n, const __nv_bfloat16* tab, const int* idx, uint32_t* sink) {
  const double gb = (double)n * D * 2 / 1e9;
  printf("== D=%d (row %d B), n=%d rows, %.2f GB gathered\n", D, D * 2, n, gb);
  for (int warps : {8, 12, 16, 24, 32}) {
    for (int split : {1}) {
      const size_t sm2 = 8192 + (size_t)warps * 2 * D * 2, sm4 = 8192 + (size_t)warps * 4 * D * 2;
      cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);
      float t2 = sm2 <= 227 * 1024 ? time_us([&] { k_bulk<2><<<148, warps * 32, sm2>>>(tab, idx, n, D, split, sink); }) : -1.f;
      float t4 = sm4 <= 227 * 1024 ? time_us([&] { k_bulk<4><<<148, warps * 32, sm4>>>(tab, idx, n, D, split, sink); }) : -1.f;
      printf("bulk   warps %2d ops/row %d: 2 stages %7.1f us %6.0f GB/s | 4 stages %7.1f us %6.0f GB/s\n", warps, split, t2,
             gb / t2 * 1e6, t4, t4 > 0 ? gb / t4 * 1e6 : 0.0);
    }
  }
  for (int warps : {8, 12, 16, 24, 32}) {
    float t2 = time_us([&] { k_ldg<2, CPL><<<148, warps * 32>>>(tab, idx, n, D, sink); });
    float t4 = time_us([&] { k_ldg<4, CPL><<<148, warps * 32>>>(tab, idx, n, D, sink); });
    printf("ldg    warps %2d: 2 rows in flight %7.1f us %6.0f GB/s | 4 rows %7.1f us %6.0f GB/s\n", warps, t2, gb / t2 * 1e6, t4,
           gb / t4 * 1e6);
  }
  for (int warps : {8, 12, 16, 24, 32}) {
    const size_t sm2 = 8192 + (size_t)warps * 2 * D * 2, sm4 = 8192 + (size_t)warps * 4 * D * 2;
    cudaFuncSetAttribute(k_cpasync<2, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    cudaFuncSetAttribute(k_cpasync<4, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);
    float t2 = sm2 <= 227 * 1024 ? time_us([&] { k_cpasync<2, CPL><<<148, warps * 32, sm2>>>(tab, idx, n, D, sink); }) : -1.f;
    float t4 = sm4 <= 227 * 1024 ? time_us([&] { k_cpasync<4, CPL><<<148, warps * 32, sm4>>>(tab, idx, n, D, sink); }) : -1.f;
    printf("ldgsts warps %2d: 2 stages %7.1f us %6.0f GB/s | 4 stages %7.1f us %6.0f GB/s\n", warps, t2, gb / t2 * 1e6, t4,
           t4 > 0 ? gb / t4 * 1e6 : 0.0);
  }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1 << 20;
  const int V = 400000;  // 400 K rows: 614 MB / 819 MB tables, far larger than the 126 MB L2
  __nv_bfloat16* tab;
  int* idx;
  uint32_t* sink;
  cudaMalloc(&tab, (size_t)V * 1024 * 2);
  cudaMemset(tab, 1, (size_t)V * 1024 * 2);
  cudaMalloc(&idx, (size_t)n * 4);
  cudaMalloc(&sink, 4);
  std::vector<int> h(n);
  uint32_t s = 12345;
  for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % V; }
  cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
  run<3>(768, n, tab, idx, sink);
  run<4>(1024, n, tab, idx, sink);
  return 0;
}
