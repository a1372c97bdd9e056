// Micro-benchmark: how fast can one B200 accumulate N*Do fp32 values into a small [Vb, bd] table?
// Brackets the byte-gradient scatter-add of mot_bwd_kernel (design evidence, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu && ./red_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

constexpr int Vb = 458, bd = 48, bpt = 16, Do = 768, NREP = 16;

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__device__ __forceinline__ void red4(float* a, float x) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a), "f"(x) : "memory");
}
__device__ __forceinline__ void red2(float* a, float x) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(a), "f"(x) : "memory");
}
__device__ __forceinline__ void red1(float* a, float x) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a), "f"(x) : "memory");
}

// mode 0: lane owns 8 consecutive floats, two v4 REDs (current kernel)
// mode 1: lane owns 4 consecutive floats per pass (contiguous 512 B per warp instruction), v4
// mode 2: v2, contiguous 256 B per warp instruction
// mode 3: scalar, contiguous 128 B per warp instruction
template <int MODE>
__global__ void __launch_bounds__(384, 1) k_red(float* acc, int n_pos, int nrep) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gw = blockIdx.x * nw + warp, W = gridDim.x * nw;
  float* accp = acc + (size_t)(gw % nrep) * Vb * bd;
  uint32_t seed = 1234567u + gw * 7919u;
  for (int pos = gw; pos < n_pos; pos += W) {
    // 16 random ids of this position (same in every lane)
    uint32_t s = seed + pos * 31u;
    int ids[bpt];
#pragma unroll
    for (int k = 0; k < bpt; ++k) ids[k] = lcg(s) % Vb;
    const float x = (float)(pos & 7);
    if (MODE == 0) {
#pragma unroll
      for (int it = 0; it < Do / 256; ++it) {
        const int e = (it * 32 + lane) * 8, slot = e / bd, boff = e - slot * bd;
        int id = 0;
#pragma unroll
        for (int k = 0; k < bpt; ++k) if (k == slot) id = ids[k];
        float* a = accp + id * bd + boff;
        red4(a, x); red4(a + 4, x);
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int it = 0; it < Do / 128; ++it) {
        const int e = (it * 32 + lane) * 4, slot = e / bd, boff = e - slot * bd;
        int id = 0;
#pragma unroll
        for (int k = 0; k < bpt; ++k) if (k == slot) id = ids[k];
        red4(accp + id * bd + boff, x);
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int it = 0; it < Do / 64; ++it) {
        const int e = (it * 32 + lane) * 2, slot = e / bd, boff = e - slot * bd;
        int id = 0;
#pragma unroll
        for (int k = 0; k < bpt; ++k) if (k == slot) id = ids[k];
        red2(accp + id * bd + boff, x);
      }
    } else {
#pragma unroll
      for (int it = 0; it < Do / 32; ++it) {
        const int e = (it * 32 + lane), slot = e / bd, boff = e - slot * bd;
        int id = 0;
#pragma unroll
        for (int k = 0; k < bpt; ++k) if (k == slot) id = ids[k];
        red1(accp + id * bd + boff, x);
      }
    }
  }
}

// mode S: shared-memory fp32 accumulator per CTA (atomicAdd on shared), flushed with REDs at the end
__global__ void __launch_bounds__(384, 1) k_smem(float* acc, int n_pos) {
  extern __shared__ float sacc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gw = blockIdx.x * nw + warp, W = gridDim.x * nw;
  for (int i = threadIdx.x; i < Vb * bd; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  uint32_t seed = 1234567u + gw * 7919u;
  for (int pos = gw; pos < n_pos; pos += W) {
    uint32_t s = seed + pos * 31u;
    int ids[bpt];
#pragma unroll
    for (int k = 0; k < bpt; ++k) ids[k] = lcg(s) % Vb;
    const float x = (float)(pos & 7);
#pragma unroll
    for (int it = 0; it < Do / 32; ++it) {
      const int e = (it * 32 + lane), slot = e / bd, boff = e - slot * bd;
      int id = 0;
#pragma unroll
      for (int k = 0; k < bpt; ++k) if (k == slot) id = ids[k];
      atomicAdd(&sacc[id * bd + boff], x);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Vb * bd; i += blockDim.x) red1(acc + i, sacc[i]);
}

// mode T: TMA bulk reduce (cp.reduce.async.bulk ... add.f32) of one 192-byte row per (position, slot)
__global__ void __launch_bounds__(384, 1) k_bulk(float* acc, int n_pos, int nrep) {
  extern __shared__ __align__(128) float stage[];  // per warp: 2 x Do floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gw = blockIdx.x * nw + warp, W = gridDim.x * nw;
  float* accp = acc + (size_t)(gw % nrep) * Vb * bd;
  float* my = stage + (size_t)warp * 2 * Do;
  uint32_t seed = 1234567u + gw * 7919u;
  int buf = 0;
  for (int pos = gw; pos < n_pos; pos += W) {
    uint32_t s = seed + pos * 31u;
    int ids[bpt];
#pragma unroll
    for (int k = 0; k < bpt; ++k) ids[k] = lcg(s) % Vb;
    const float x = (float)(pos & 7);
    float* st = my + buf * Do;
    // the previous bulk group that read this buffer must have finished reading it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int it = 0; it < Do / 128; ++it) *reinterpret_cast<float4*>(st + (it * 32 + lane) * 4) = make_float4(x, x, x, x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < bpt) {
      int id = 0;
#pragma unroll
      for (int k = 0; k < bpt; ++k) if (k == lane) id = ids[k];
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(st + lane * bd);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(accp + id * bd), "r"(src),
                   "r"(bd * 4)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float time_it(F f, int reps = 20) {
  for (int i = 0; i < 3; ++i) f();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  if (cudaGetLastError() != cudaSuccess) printf("CUDA error\n");
  return ms / reps * 1e3f;
}

int main(int argc, char** argv) {
  const int n_pos = argc > 1 ? atoi(argv[1]) : 49152;
  float* acc;
  cudaMalloc(&acc, (size_t)NREP * Vb * bd * 4 * 4);
  cudaMemset(acc, 0, (size_t)NREP * Vb * bd * 4 * 4);
  int sms = 148;
  printf("n_pos %d, %d floats per position -> %.1f M fp32 adds\n", n_pos, Do, n_pos * (double)Do / 1e6);
  for (int nrep : {1, 4, 16}) {
    printf("nrep %2d: v4 8-per-lane %.1f us | v4 contiguous %.1f us | v2 %.1f us | scalar %.1f us\n", nrep,
           time_it([&] { k_red<0><<<sms, 384>>>(acc, n_pos, nrep); }), time_it([&] { k_red<1><<<sms, 384>>>(acc, n_pos, nrep); }),
           time_it([&] { k_red<2><<<sms, 384>>>(acc, n_pos, nrep); }), time_it([&] { k_red<3><<<sms, 384>>>(acc, n_pos, nrep); }));
  }
  for (int thr : {384, 768, 1024}) {
    printf("threads %d nrep 16: v4 8-per-lane %.1f us | v4 contiguous %.1f us\n", thr,
           time_it([&] { k_red<0><<<sms, thr>>>(acc, n_pos, 16); }), time_it([&] { k_red<1><<<sms, thr>>>(acc, n_pos, 16); }));
  }
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, Vb * bd * 4);
  printf("shared atomicAdd(float) per CTA: %.1f us\n", time_it([&] { k_smem<<<sms, 384, Vb * bd * 4>>>(acc, n_pos); }));
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 2 * Do * 4);
  for (int nrep : {1, 16})
    printf("TMA bulk reduce add.f32 (192 B per op) nrep %d: %.1f us\n", nrep,
           time_it([&] { k_bulk<<<sms, 384, 12 * 2 * Do * 4>>>(acc, n_pos, nrep); }));
  return 0;
}
