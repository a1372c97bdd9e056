// Dense projection of the concat+projection mixin variants on the 5th-generation tensor cores (tcgen05).
//
// Replaces (reference, read-only at /root/reference in the build container):
//   forward : `F.linear(x, W)` of mixin_bytes (runs/7:233-234) and ByteMixinConcat / CastedLinear
//             (spt/train_gpt.py:443,185-186): Y[n, Do] = X[n, K] . W[Do, K]^T, X = [tok | bytes] from mot_embed_fwd
//   backward: autograd of that line: dX = dY . W  and  dW = dY^T . X
//
// One persistent, warp-specialised kernel (sm_100a): warp 0 = TMA producer (cp.async.bulk.tensor, 128-byte swizzle),
// warp 1 = one elected thread issuing tcgen05.mma (128 x 256 x 16, bf16 in, fp32 accumulators in TMEM, double
// buffered: 2 x 256 of the 512 TMEM columns), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> registers
// -> global).  Operands are read in place in either storage order through the UMMA shared-memory descriptors:
//   K-major  : contraction dimension contiguous      (X and W in the forward)
//   MN-major : contraction dimension strided         (W in dX; dY and X in dW -- no transposed copies are made)
// dW contracts over the tokens, so there are few output tiles: the token range is split over CTAs and the fp32
// partial tiles are reduced with red.global.add.v4.f32 into a zeroed fp32 buffer.
#include <cuda.h>

#include "mot_common.cuh"

namespace mot {

constexpr int kGemmThreads = 256;
// CTA tile 128 x 256 x (128 bytes of contraction): one 128-byte swizzle row = 64 bf16 or 32 fp32 (tf32) elements, and
// one tcgen05.mma consumes 32 bytes of it (K = 16 for kind::f16, K = 8 for kind::tf32), so both element types share
// the stage layout, the byte offsets of the descriptors and four MMAs per stage.
constexpr int BM = 128, BN = 256;
constexpr int kRowBytes = 128, kMmaBytes = 32;
constexpr int kGemmStages = 4;
constexpr uint32_t kABytes = BM * kRowBytes, kBBytes = BN * kRowBytes, kStageBytes = kABytes + kBBytes;
constexpr uint32_t kTmemCols = 512;  // two fp32 accumulators of BN columns

enum { EPI_STORE_BF16 = 0, EPI_STORE_F32 = 1, EPI_RED_F32 = 2 };

struct GemmParams {
  void* C;
  const float* bias;  // [N] fp32 or null (added in the epilogue)
  long long ldc;
  int M, N, K;
  int m_tiles, n_tiles, k_blocks;  // k_blocks = ceil(K / elements per 128-byte row)
  int splits;                      // split of the contraction range (EPI_RED_F32 only)
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] . B[smem]
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of the warp's TMEM quarter
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets (all >> 4),
// version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
//   K-major  tile [rows][64 elements]: rows at 128 B pitch, 8-row swizzle atoms 1024 B apart (SBO); LBO unused (1)
//   MN-major tile [64-element MN block][k rows at 128 B pitch]: 8-k atoms 1024 B apart (SBO), MN blocks LBO apart
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;  // version
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, operand format (1 = bf16, 2 = tf32), operand
// majors, N, M
__host__ __device__ constexpr uint32_t make_idesc(bool tf32, bool a_mn, bool b_mn, int m, int n) {
  return (1u << 4) | ((tf32 ? 2u : 1u) << 7) | ((tf32 ? 2u : 1u) << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct __align__(8) GemmBarriers {
  uint64_t full[kGemmStages], empty[kGemmStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// TE: operand element type (__nv_bfloat16 -> kind::f16, float -> kind::tf32, fp32 values read in place)
template <typename TE, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
mot_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned stage buffers (128-byte swizzle atoms are 1024 B), barriers behind them
  unsigned char* stage_base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(stage_base + (size_t)kGemmStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = lane_id();
  constexpr bool kTf32 = sizeof(TE) == 4;
  constexpr int KE = kRowBytes / sizeof(TE);     // contraction elements per stage (one 128-byte swizzle row): 64 / 32
  constexpr int kMnBlock = KE * kRowBytes;       // MN-major operand: one block = KE contraction rows x 128 B of MN elements

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->acc_full[a], 1);
      mbar_init(&bars->acc_empty[a], 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  } else if (warp == 2) {
    tmem_alloc(&bars->tmem_base, kTmemCols);
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // nothing above touches global memory
  const uint32_t tmem_base = bars->tmem_base;

  const int n_work = p.m_tiles * p.n_tiles * p.splits;
  const int kb_per_split = (p.k_blocks + p.splits - 1) / p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int tile = w / p.splits, split = w - tile * p.splits;
        const int m_blk = tile / p.n_tiles, n_blk = tile - m_blk * p.n_tiles;
        const int kb0 = split * kb_per_split, kb1 = min(kb0 + kb_per_split, p.k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          unsigned char* sa = stage_base + (size_t)stage * kStageBytes;
          unsigned char* sb = sa + kABytes;
          mbar_expect_tx(&bars->full[stage], kStageBytes);
          if (A_MN) {  // global [K rows][M cols]: one box of KE M-elements (128 B) x KE k-rows per MN block
#pragma unroll
            for (int j = 0; j < BM / KE; ++j) tma_load_2d(sa + j * kMnBlock, &tmap_a, m_blk * BM + j * KE, kb * KE, &bars->full[stage]);
          } else {     // global [M rows][K cols]: one box of KE k-elements (128 B) x BM rows
            tma_load_2d(sa, &tmap_a, kb * KE, m_blk * BM, &bars->full[stage]);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / KE; ++j) tma_load_2d(sb + j * kMnBlock, &tmap_b, n_blk * BN + j * KE, kb * KE, &bars->full[stage]);
          } else {
            tma_load_2d(sb, &tmap_b, kb * KE, n_blk * BN, &bars->full[stage]);
          }
          if (++stage == kGemmStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kTf32, A_MN, B_MN, BM, BN);
      constexpr int kMmaRows = kMmaBytes / sizeof(TE);  // contraction rows per MMA of an MN-major operand: 16 / 8
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int tile = w / p.splits, split = w - tile * p.splits;
        const int kb0 = split * kb_per_split, kb1 = min(kb0 + kb_per_split, p.k_blocks);
        (void)tile;
        mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)stage * kStageBytes), sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kRowBytes / kMmaBytes; ++k) {
            // advance along the contraction: 32 B inside the swizzle row (K-major) / 16 or 8 k-rows of 128 B (MN-major)
            const uint64_t da = A_MN ? make_desc(sa + k * (kMmaRows * kRowBytes), kMnBlock, 1024) : make_desc(sa + k * kMmaBytes, 16, 1024);
            const uint64_t db = B_MN ? make_desc(sb + k * (kMmaRows * kRowBytes), kMnBlock, 1024) : make_desc(sb + k * kMmaBytes, 16, 1024);
            umma<kTf32>(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&bars->empty[stage]);  // frees the stage once these MMAs have read it
          if (++stage == kGemmStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&bars->acc_full[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;  // this warp reads TMEM lanes 32q .. 32q+31 = tile rows
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int tile = w / p.splits;
      const int m_blk = tile / p.n_tiles, n_blk = tile - m_blk * p.n_tiles;
      const int kb0 = (w - tile * p.splits) * kb_per_split;
      const bool has_work = kb0 < p.k_blocks;  // an empty split leaves garbage in TMEM: skip its stores
      mbar_wait(&bars->acc_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_row + c, v);
        tmem_ld_wait();
        const int col = n_blk * BN + c;
        if (row < p.M && col < p.N && has_work) {
          const int nvalid = min(32, p.N - col);  // N is a multiple of 8 (validated on the host)
          if (EPI == EPI_STORE_BF16) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (j < nvalid) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]) + (p.bias ? ld_g(p.bias + col + j + e) : 0.f);
                Vec<__nv_bfloat16, 8>::stg(dst + j, f);
              }
            }
          } else {
            float* dst = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < nvalid) {
                float f[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) f[e] = __uint_as_float(v[j + e]) + ((EPI == EPI_STORE_F32 && p.bias) ? ld_g(p.bias + col + j + e) : 0.f);
                if (EPI == EPI_RED_F32) atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(f[0], f[1], f[2], f[3]));
                else Vec<float, 4>::stg(dst + j, f);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ======================================================================================
// Row-wise rms_norm over the projected rows (norm(F.linear(...)), runs/7:234, spt/train_gpt.py:443): fp32 math on the
// bf16 / fp32 rows, one warp per row, 16-byte accesses.  HBM bound: 2 * D * e bytes per row forward, 3 * D * e backward.
// ======================================================================================
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_fwd_kernel(const T* y, T* out, long long n, int D, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = lane_id();
  const long long W = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += W) {
    const T* src = y + (size_t)r * D;
    float ss = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float v[8];
      Vec<T, 8>::unpack(Vec<T, 8>::ldg_raw(src + c), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
    }
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss / (float)D + eps);
    for (int c = lane * 8; c < D; c += 256) {
      float v[8];
      Vec<T, 8>::unpack(Vec<T, 8>::ldg_raw(src + c), v);  // second read hits L1/L2
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= rs;
      Vec<T, 8>::stg(out + (size_t)r * D + c, v);
    }
  }
}

// dy = rs * g - y * rs^3 * mean(g . y)
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const T* y, const T* g, T* dy, long long n,
                                                          int D, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = lane_id();
  const long long W = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += W) {
    const T* ys = y + (size_t)r * D;
    const T* gs = g + (size_t)r * D;
    float ss = 0.f, gy = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float a[8], b[8];
      Vec<T, 8>::unpack(Vec<T, 8>::ldg_raw(ys + c), a);
      Vec<T, 8>::unpack(Vec<T, 8>::ldg_raw(gs + c), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        ss += a[e] * a[e];
        gy += a[e] * b[e];
      }
    }
    warp_sum2(ss, gy);
    const float rs = rsqrtf(ss / (float)D + eps);
    const float coef = rs * rs * rs * gy / (float)D;
    for (int c = lane * 8; c < D; c += 256) {
      float a[8], b[8], o[8];
      Vec<T, 8>::unpack(Vec<T, 8>::ldg_raw(ys + c), a);
      Vec<T, 8>::unpack(Vec<T, 8>::ldg_raw(gs + c), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rs * b[e] - coef * a[e];
      Vec<T, 8>::stg(dy + (size_t)r * D + c, o);
    }
  }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* src, __nv_bfloat16* dst, long long n8) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    Vec<float, 8>::unpack(Vec<float, 8>::ldg_raw(src + i * 8), v);
    Vec<__nv_bfloat16, 8>::stg(dst + i * 8, v);
  }
}

// dst[c][r] = src[r][c] for fp32 (32 x 32 tiles through shared memory).  Only the fp32 (TF32) backward uses it: the
// hardware's MN-major operand layout for 32-bit elements is a different swizzle atom than the 16-bit one this kernel
// stages with TMA, so the two small fp32 backward GEMMs (mathblations: 11 K tokens) run on transposed copies instead.
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* src, float* dst, int R, int C, long long ld_dst) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (r0 + j < R && c0 + tx < C) tile[j][tx] = src[(size_t)(r0 + j) * C + c0 + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (c0 + j < C && r0 + tx < R) dst[(size_t)(c0 + j) * ld_dst + r0 + tx] = tile[tx][j];
}

// ======================================================================================
// Host side
// ======================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// matrix [rows][cols] row-major (cols contiguous) of bf16 (esz 2) or fp32 (esz 4); box = box_cols x box_rows, 128-byte
// swizzle, zero fill out of bounds
static int make_tmap(CUtensorMap* m, const void* base, long long rows, long long cols, int box_cols, int box_rows, int esz,
                     long long ld = 0) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MOT_ERR_CUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)(ld > 0 ? ld : cols) * esz};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MOT_OK : MOT_ERR_BAD_ARG;
}

// C[M,N] (+)= op(A) . op(B)^T with the contraction dimension of length K.
//   a_mn == false: A is [M][K] row-major;  true: A is [K][M] row-major.  Same for B with N.
template <typename TE, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, long long ldc, int splits,
                       cudaStream_t s, long long lda = 0, long long ldb = 0) {
  int sms = 0, optin = 0;
  if (int rc = device_props(&sms, &optin)) return rc;
  constexpr int ES = sizeof(TE), KE = kRowBytes / ES;
  CUtensorMap ta, tb;
  if (int rc = A_MN ? make_tmap(&ta, A, K, M, KE, KE, ES, lda) : make_tmap(&ta, A, M, K, KE, BM, ES, lda)) return rc;
  if (int rc = B_MN ? make_tmap(&tb, B, K, N, KE, KE, ES, ldb) : make_tmap(&tb, B, N, K, KE, BN, ES, ldb)) return rc;
  GemmParams p{};
  p.C = C; p.bias = bias; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.k_blocks = (K + KE - 1) / KE;
  p.splits = splits < 1 ? 1 : splits;
  const size_t smem = (size_t)kGemmStages * kStageBytes + sizeof(GemmBarriers) + 1024;
  auto kern = mot_gemm_kernel<TE, A_MN, B_MN, EPI>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return check_launch();
  long long work = (long long)p.m_tiles * p.n_tiles * p.splits;
  const int grid = (int)(work < sms ? work : sms);
  launch_pdl(kern, dim3(grid), dim3(kGemmThreads), smem, s, ta, tb, p);
  count_launch();
  return check_launch();
}

static bool ok16(const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static void launch_transpose(const float* src, float* dst, int R, int C, long long ld_dst, cudaStream_t s) {
  launch_pdl(transpose_f32_kernel, dim3((C + 31) / 32, (R + 31) / 32), dim3(256), 0, s, src, dst, R, C, ld_dst);
  count_launch();
}
static long long pad4(long long n) { return (n + 3) / 4 * 4; }

}  // namespace mot

using namespace mot;

extern "C" int mot_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t n_tokens, int32_t in_dim,
                              int32_t out_dim, int32_t dtype, int32_t y_f32, void* stream) {
  if (n_tokens < 0 || n_tokens > 0x7fffffffLL || in_dim <= 0 || out_dim <= 0) return MOT_ERR_BAD_ARG;
  if (dtype != MOT_BF16 && dtype != MOT_F32) return MOT_ERR_UNSUPPORTED;
  if (dtype == MOT_F32 && !y_f32) return MOT_ERR_UNSUPPORTED;
  if (in_dim % 8 || out_dim % 8) return MOT_ERR_MISALIGNED;
  if (n_tokens == 0) return MOT_OK;
  if (!ok16(x) || !ok16(w) || !ok16(y)) return x && w && y ? MOT_ERR_MISALIGNED : MOT_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int n = (int)n_tokens;
  if (dtype == MOT_F32) return launch_gemm<float, false, false, EPI_STORE_F32>(x, w, y, bias, n, out_dim, in_dim, out_dim, 1, s);
  return y_f32 ? launch_gemm<__nv_bfloat16, false, false, EPI_STORE_F32>(x, w, y, bias, n, out_dim, in_dim, out_dim, 1, s)
               : launch_gemm<__nv_bfloat16, false, false, EPI_STORE_BF16>(x, w, y, bias, n, out_dim, in_dim, out_dim, 1, s);
}

extern "C" size_t mot_linear_workspace_bytes(int64_t n_tokens, int32_t in_dim, int32_t out_dim, int32_t dtype) {
  if (dtype != MOT_F32 || n_tokens <= 0 || in_dim <= 0 || out_dim <= 0) return 0;  // bf16 operands are read in place
  const size_t wt = (size_t)in_dim * out_dim * 4;
  const size_t acts = (size_t)pad4(n_tokens) * ((size_t)in_dim + out_dim) * 4;
  return (wt > acts ? wt : acts) + 256;
}

extern "C" int mot_linear_bwd_input(const void* dy, const void* w, void* dx, int64_t n_tokens, int32_t in_dim, int32_t out_dim,
                                    int32_t dtype, void* workspace, size_t ws_bytes, void* stream) {
  if (n_tokens < 0 || n_tokens > 0x7fffffffLL || in_dim <= 0 || out_dim <= 0) return MOT_ERR_BAD_ARG;
  if (dtype != MOT_BF16 && dtype != MOT_F32) return MOT_ERR_UNSUPPORTED;
  if (in_dim % 8 || out_dim % 8) return MOT_ERR_MISALIGNED;
  if (n_tokens == 0) return MOT_OK;
  if (!ok16(dy) || !ok16(w) || !ok16(dx)) return dy && w && dx ? MOT_ERR_MISALIGNED : MOT_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == MOT_F32) {  // dX = dY . (W^T)^T with W^T [K, Do] built in the workspace
    if (!workspace || ws_bytes < mot_linear_workspace_bytes(n_tokens, in_dim, out_dim, dtype)) return MOT_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return MOT_ERR_MISALIGNED;
    float* wt = reinterpret_cast<float*>(workspace);
    launch_transpose(reinterpret_cast<const float*>(w), wt, out_dim, in_dim, out_dim, s);
    return launch_gemm<float, false, false, EPI_STORE_F32>(dy, wt, dx, nullptr, (int)n_tokens, in_dim, out_dim, in_dim, 1, s);
  }
  // dX[n, K] = dY[n, Do] . W[Do, K]: contraction over Do; dY is K-major, W is read in place as an MN-major operand
  return launch_gemm<__nv_bfloat16, false, true, EPI_STORE_BF16>(dy, w, dx, nullptr, (int)n_tokens, in_dim, out_dim, in_dim, 1, s);
}

extern "C" int mot_linear_bwd_weight(const void* dy, const void* x, float* dw_f32, void* dw_bf16, int64_t n_tokens, int32_t in_dim,
                                     int32_t out_dim, int32_t dtype, void* workspace, size_t ws_bytes, void* stream) {
  if (n_tokens < 0 || n_tokens > 0x7fffffffLL || in_dim <= 0 || out_dim <= 0) return MOT_ERR_BAD_ARG;
  if (dtype != MOT_BF16 && dtype != MOT_F32) return MOT_ERR_UNSUPPORTED;
  if (in_dim % 8 || out_dim % 8) return MOT_ERR_MISALIGNED;
  if (!dw_f32 || (reinterpret_cast<uintptr_t>(dw_f32) & 15u)) return MOT_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t n_el = (size_t)out_dim * in_dim;
  if (dtype == MOT_F32 && n_tokens > 0) {  // dW = (dY^T) . (X^T)^T on transposed copies [Do, n] and [K, n] in the workspace
    if (!ok16(dy) || !ok16(x)) return dy && x ? MOT_ERR_MISALIGNED : MOT_ERR_BAD_ARG;
    if (!workspace || ws_bytes < mot_linear_workspace_bytes(n_tokens, in_dim, out_dim, dtype)) return MOT_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return MOT_ERR_MISALIGNED;
    const long long ld = pad4(n_tokens);
    float* dyT = reinterpret_cast<float*>(workspace);
    float* xT = dyT + (size_t)out_dim * ld;
    launch_transpose(reinterpret_cast<const float*>(dy), dyT, (int)n_tokens, out_dim, ld, s);
    launch_transpose(reinterpret_cast<const float*>(x), xT, (int)n_tokens, in_dim, ld, s);
    return launch_gemm<float, false, false, EPI_STORE_F32>(dyT, xT, dw_f32, nullptr, out_dim, in_dim, (int)n_tokens, in_dim, 1, s, ld, ld);
  }
  if (cudaMemsetAsync(dw_f32, 0, n_el * 4, s) != cudaSuccess) return check_launch();
  if (n_tokens > 0) {
    if (!ok16(dy) || !ok16(x)) return dy && x ? MOT_ERR_MISALIGNED : MOT_ERR_BAD_ARG;
    int sms = 0, optin = 0;
    if (int rc = device_props(&sms, &optin)) return rc;
    // dW[Do, K] = dY^T . X: contraction over the tokens, both operands MN-major; split the token range so that the
    // tiles x splits fill the machine (at least 4 k-blocks per split)
    const int tiles = ((out_dim + BM - 1) / BM) * ((in_dim + BN - 1) / BN);
    const int ke = dtype == MOT_F32 ? kRowBytes / 4 : kRowBytes / 2;
    const long long kblocks = (n_tokens + ke - 1) / ke;
    // pick the split count (<= 16, >= 4 k-blocks each) whose work items fill whole waves of the persistent grid best
    long long splits = 1;
    double best = 0.0;
    for (long long sp = 1; sp <= 16 && sp <= (kblocks + 3) / 4; ++sp) {
      const long long work = (long long)tiles * sp, waves = (work + sms - 1) / sms;
      const double eff = (double)work / (double)(waves * sms);
      if (eff > best + 0.02) {
        best = eff;
        splits = sp;
      }
    }
    if (int rc = launch_gemm<__nv_bfloat16, true, true, EPI_RED_F32>(dy, x, dw_f32, nullptr, out_dim, in_dim, (int)n_tokens, in_dim,
                                                                     (int)splits, s))
      return rc;
  }
  if (dw_bf16) {  // the runs keep the mixin weight in bf16 (runs/7:249): cast the reduced gradient once
    if (reinterpret_cast<uintptr_t>(dw_bf16) & 15u) return MOT_ERR_MISALIGNED;
    const long long n8 = (long long)(n_el / 8);
    long long blocks = (n8 + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    launch_pdl(f32_to_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, s, (const float*)dw_f32, reinterpret_cast<__nv_bfloat16*>(dw_bf16), n8);
    count_launch();
  }
  return check_launch();
}

extern "C" int mot_rmsnorm_fwd(const void* y, void* out, int64_t n_rows, int32_t dim, int32_t dtype, float eps, void* stream) {
  if (n_rows < 0 || dim <= 0) return MOT_ERR_BAD_ARG;
  if (dim % 8) return MOT_ERR_MISALIGNED;
  if (dtype != MOT_BF16 && dtype != MOT_F32) return MOT_ERR_UNSUPPORTED;
  if (n_rows == 0) return MOT_OK;
  if (!ok16(y) || !ok16(out)) return y && out ? MOT_ERR_MISALIGNED : MOT_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  long long blocks = (n_rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dtype == MOT_BF16)
    launch_pdl(rmsnorm_fwd_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(256), 0, s, (const __nv_bfloat16*)y, (__nv_bfloat16*)out,
               (long long)n_rows, (int)dim, eps);
  else
    launch_pdl(rmsnorm_fwd_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, s, (const float*)y, (float*)out, (long long)n_rows, (int)dim, eps);
  count_launch();
  return check_launch();
}

extern "C" int mot_rmsnorm_bwd(const void* y, const void* grad_out, void* dy, int64_t n_rows, int32_t dim, int32_t dtype, float eps,
                               void* stream) {
  if (n_rows < 0 || dim <= 0) return MOT_ERR_BAD_ARG;
  if (dim % 8) return MOT_ERR_MISALIGNED;
  if (dtype != MOT_BF16 && dtype != MOT_F32) return MOT_ERR_UNSUPPORTED;
  if (n_rows == 0) return MOT_OK;
  if (!ok16(y) || !ok16(grad_out) || !ok16(dy)) return y && grad_out && dy ? MOT_ERR_MISALIGNED : MOT_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  long long blocks = (n_rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dtype == MOT_BF16)
    launch_pdl(rmsnorm_bwd_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(256), 0, s, (const __nv_bfloat16*)y,
               (const __nv_bfloat16*)grad_out, (__nv_bfloat16*)dy, (long long)n_rows, (int)dim, eps);
  else
    launch_pdl(rmsnorm_bwd_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, s, (const float*)y, (const float*)grad_out, (float*)dy,
               (long long)n_rows, (int)dim, eps);
  count_launch();
  return check_launch();
}
