// Fused byte-mix embedding forward / backward for sm_100a.
//
// Replaces (reference, read-only at /root/reference in the build container):
//   forward : runs/71:312-314 + mixin_bytes :228-230 and the other gather/pool/add/norm variants
//             of SURVEY.md 2.4; spt/train_gpt.py:342-379 (+ the cat of :443 as GEMM A operand)
//   backward: the autograd graph of those lines: rms_norm backward, split, and two
//             embedding_dense_backward scatter-adds producing DENSE [V,Dt] / [Vb,bd] grads.
//
// Data layout in HBM: E_tok [V,Dt], E_byte [Vb,bd], out / grad_out [N,Do] row-major, element type T
// (bf16 or fp32); token ids int32 [N]; byte ids int32/int64, token-major [N,bpt] or slot-major [bpt,N];
// ttb table [V,bpt] int16 (or the reference's fp32 / bf16 float containers).
//
// One warp owns one position (forward) or one token row with its occurrences (backward); a lane owns
// chunks of 8 consecutive elements: chunk c = it*32 + lane, so a warp-wide access is one contiguous
// 512 B (bf16) segment per `it`.  The byte table lives in shared memory (staged once per CTA with a
// bulk async copy), so the only HBM streams are token rows, grad rows, output rows and the dense grad.
#pragma once
#include <cstdlib>
#include "mot_common.cuh"

namespace mot {

constexpr int kFwdThreads = 512;
constexpr int kBwdThreads = 384;

struct EmbedParams {
  const int32_t* tok;
  const void* ids;
  const void* ttb;
  const void* E_tok;
  const void* E_byte;
  const float* lam;
  void* out;
  // backward
  const void* gout;
  void* gE_tok;
  void* gE_byte;
  float* g_lam;
  // plan / workspace views
  int* cnt;            // [V]     histogram, doubles as fill cursor
  int* off;            // [V+1]   exclusive scan of cnt
  int* item_off;       // [V+1]   exclusive scan of ceil(cnt/L)
  int* hot_off;        // [V+1]   exclusive scan of [cnt > L]
  int* pslot_off;      // [V+1]   exclusive scan of [cnt > L] * ceil(cnt/L)
  int* order;          // [N]     positions grouped by token id
  int4* items;         // [max_items]  {v, start, cnt, partial slot or -1}
  int4* hot_rows;      // [max_hot]    {v, first partial slot, n chunks, 0}
  float* partial;      // [max_pslots, Dt] fp32 partial sums of hot rows
  float* byte_acc;     // [Vb*bd] fp32
  float* lam_acc;      // [2]
  long long N, T;
  int V, Vb, bpt, Dt, bd, Do, combine, flags, ttb_dtype;
  int n_chunks;  // Do / 8
  int L;         // occurrences per work item
  int acc_stride;  // floats per row of the shared-memory accumulators (bd + 2: spreads rows over banks)
  int tab_smem;  // 1: byte table staged in shared memory; 0: too large, rows read through L1/L2
  float eps;
};

struct ChunkMap {
  int toff;  // element offset in the token row, -1: chunk has no token part
  int slot;  // byte slot, -1: no byte part, -2: mean over all slots
  int boff;  // element offset inside the byte row
};

__device__ __forceinline__ ChunkMap chunk_map(const EmbedParams& p, int c) {
  ChunkMap m{-1, -1, 0};
  if (c >= p.n_chunks) return m;
  const int e = c * kChunk;
  switch (p.combine) {
    case MOT_ADD:
      m.toff = e;
      m.slot = e / p.bd;
      m.boff = e - m.slot * p.bd;
      break;
    case MOT_CONCAT: {
      const int db = p.bpt * p.bd;
      int eb = -1;
      if (p.flags & MOT_F_BYTES_FIRST) {
        if (e < db) eb = e; else m.toff = e - db;
      } else {
        if (e < p.Dt) m.toff = e; else eb = e - p.Dt;
      }
      if (eb >= 0) {
        m.slot = eb / p.bd;
        m.boff = eb - m.slot * p.bd;
      }
    } break;
    case MOT_TOK_ONLY:
      m.toff = e;
      break;
    case MOT_BYTES_ONLY:
      m.slot = e / p.bd;
      m.boff = e - m.slot * p.bd;
      break;
    case MOT_MEAN:
      m.toff = e;
      m.slot = -2;
      m.boff = e;
      break;
  }
  return m;
}

__device__ __forceinline__ int clampi(int v, int hi) { return min(max(v, 0), hi); }

// byte id of (position, slot); slot < bpt.  Out-of-range ids are clamped (the reference device-asserts).
__device__ __forceinline__ int fetch_id(const EmbedParams& p, long long pos, int slot) {
  int id;
  if (p.flags & MOT_F_IDS_FROM_TTB) {
    long long tokpos = pos;
    int k = slot;
    if (p.flags & MOT_F_TTB_SCRAMBLE) {  // runs/71:479: flat byte i*T + s of the row
      const long long row = pos / p.T, s = pos - row * p.T;
      const long long f = (long long)slot * p.T + s;
      tokpos = row * p.T + f / p.bpt;
      k = (int)(f % p.bpt);
    }
    const int tv = clampi(__ldg(p.tok + tokpos), p.V - 1);
    const size_t e = (size_t)tv * p.bpt + k;
    if (p.ttb_dtype == MOT_TTB_I16) id = __ldg(reinterpret_cast<const short*>(p.ttb) + e);
    else if (p.ttb_dtype == MOT_TTB_F32) id = (int)__ldg(reinterpret_cast<const float*>(p.ttb) + e);
    else id = (int)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.ttb)[e]);
  } else {
    const long long idx = (p.flags & MOT_F_SLOT_MAJOR) ? (long long)slot * p.N + pos : pos * p.bpt + slot;
    id = (p.flags & MOT_F_IDS_I64) ? (int)__ldg(reinterpret_cast<const long long*>(p.ids) + idx)
                                   : __ldg(reinterpret_cast<const int*>(p.ids) + idx);
  }
  return clampi(id, p.Vb - 1);
}

// Stage E_byte into shared memory with one bulk async copy per 32 KB and compute the per-row
// rms scale (byte_norm).  rs[r] = rsqrt(mean(row^2) + eps) or 1.
template <typename T>
__device__ __forceinline__ typename Vec8<T>::Raw tab_load(const EmbedParams& p, const T* tab, size_t off) {
  return p.tab_smem ? Vec8<T>::lds_raw(tab + off) : Vec8<T>::ldg_raw(reinterpret_cast<const T*>(p.E_byte) + off);
}

template <typename T>
__device__ void stage_byte_table(const EmbedParams& p, T* tab, float* rs, uint64_t* bar) {
  if (p.tab_smem) {
    const uint32_t bytes = (uint32_t)p.Vb * p.bd * sizeof(T);
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, bytes);
      for (uint32_t done = 0; done < bytes;) {
        const uint32_t n = min(bytes - done, 32768u);
        bulk_g2s(reinterpret_cast<char*>(tab) + done, reinterpret_cast<const char*>(p.E_byte) + done, n, bar);
        done += n;
      }
    }
    mbar_wait(bar, 0);
  }
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = lane_id();
  const bool bn = (p.flags & MOT_F_BYTE_NORM) != 0;
  for (int r = warp; r < p.Vb; r += nw) {
    float ss = 0.f;
    if (bn) {
      for (int c = lane; c < p.bd / kChunk; c += 32) {
        float v[8];
        Vec8<T>::unpack(tab_load<T>(p, tab, (size_t)r * p.bd + c * kChunk), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
      }
      ss = warp_sum(ss);
    }
    if (lane == 0) rs[r] = bn ? rsqrtf(ss / (float)p.bd + p.eps) : 1.f;
  }
  __syncthreads();
}

// acc[0..8) += v[0..8).  Shared memory has no native fp32 atomic add (atomicAdd lowers to one
// load + CAS spin loop PER element, 24 serial round trips per occurrence); here the four 64-bit words
// are read once, the four CAS are issued back to back and only the losers retry.
__device__ __forceinline__ void smem_add8(float* a, const float (&v)[8]) {
  unsigned long long* a2 = reinterpret_cast<unsigned long long*>(a);
  unsigned long long old[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) old[q] = *reinterpret_cast<volatile unsigned long long*>(a2 + q);
  unsigned pending = 0xFu;
  while (pending) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (pending & (1u << q)) {
        const float lo = __uint_as_float((unsigned)old[q]) + v[2 * q];
        const float hi = __uint_as_float((unsigned)(old[q] >> 32)) + v[2 * q + 1];
        const unsigned long long nw = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
        const unsigned long long prev = atomicCAS(a2 + q, old[q], nw);
        if (prev == old[q]) pending &= ~(1u << q);
        else old[q] = prev;
      }
    }
  }
}
__device__ __forceinline__ void gmem_add8(float* a, const float (&v)[8]) {
  atomicAdd(reinterpret_cast<float4*>(a), make_float4(v[0], v[1], v[2], v[3]));      // RED.E.ADD.F32x4
  atomicAdd(reinterpret_cast<float4*>(a) + 1, make_float4(v[4], v[5], v[6], v[7]));
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ======================================================================================
// Forward
// ======================================================================================
template <typename T, int CPL>
struct FwdLoad {
  typename Vec8<T>::Raw traw[CPL];
  int idreg;
};

template <typename T, int CPL>
__global__ void __launch_bounds__(kFwdThreads) mot_fwd_kernel(const EmbedParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  float* rs = reinterpret_cast<float*>(smem_raw);
  T* tab = reinterpret_cast<T*>(smem_raw + align_up((size_t)p.Vb * sizeof(float), 128));
  const bool has_tok = p.combine != MOT_BYTES_ONLY;
  const bool has_bytes = p.combine != MOT_TOK_ONLY;
  if (has_bytes) stage_byte_table<T>(p, tab, rs, &bar);

  const int lane = lane_id();
  const int nw = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * nw + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * nw;

  ChunkMap cm[CPL];
#pragma unroll
  for (int it = 0; it < CPL; ++it) cm[it] = chunk_map(p, it * 32 + lane);

  float lam_t = 1.f, lam_b = 1.f;
  if (p.flags & MOT_F_HAS_LAMBDAS) {
    lam_t = __ldg(p.lam);
    lam_b = __ldg(p.lam + 1);
  }
  if (p.combine == MOT_MEAN) lam_b /= (float)p.bpt;
  const T* E_tok = reinterpret_cast<const T*>(p.E_tok);
  T* out = reinterpret_cast<T*>(p.out);
  const bool tok_norm = (p.flags & MOT_F_TOK_NORM) != 0;
  const bool out_norm = (p.flags & MOT_F_OUT_NORM) != 0;

  auto load = [&](long long pos, FwdLoad<T, CPL>& ld) {
    ld.idreg = (has_bytes && lane < p.bpt) ? fetch_id(p, pos, lane) : 0;
    if (has_tok) {
      const int tv = clampi(__ldg(p.tok + pos), p.V - 1);
      const T* row = E_tok + (size_t)tv * p.Dt;
#pragma unroll
      for (int it = 0; it < CPL; ++it)
        if (cm[it].toff >= 0) ld.traw[it] = Vec8<T>::ldg_raw(row + cm[it].toff);
    }
  };

  auto finish = [&](long long pos, FwdLoad<T, CPL>& ld) {
    float x[CPL][8];
    float ss_t = 0.f;
#pragma unroll
    for (int it = 0; it < CPL; ++it) {
      if (cm[it].toff >= 0) {
        Vec8<T>::unpack(ld.traw[it], x[it]);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss_t += x[it][e] * x[it][e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[it][e] = 0.f;
      }
    }
    float tscale = lam_t;
    if (tok_norm) {
      ss_t = warp_sum(ss_t);
      tscale *= rsqrtf(ss_t / (float)p.Dt + p.eps);
    }
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < CPL; ++it) {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[it][e] *= tscale;
      if (has_bytes) {
        if (p.combine == MOT_MEAN) {
          for (int k = 0; k < p.bpt; ++k) {
            const int id = __shfl_sync(0xffffffffu, ld.idreg, k);
            if (cm[it].slot == -2) {
              float b[8];
              Vec8<T>::unpack(tab_load<T>(p, tab, (size_t)id * p.bd + cm[it].boff), b);
              const float bs = lam_b * rs[id];
#pragma unroll
              for (int e = 0; e < 8; ++e) x[it][e] += bs * b[e];
            }
          }
        } else {
          const int id = __shfl_sync(0xffffffffu, ld.idreg, cm[it].slot & 31);
          if (cm[it].slot >= 0) {
            float b[8];
            Vec8<T>::unpack(tab_load<T>(p, tab, (size_t)id * p.bd + cm[it].boff), b);
            const float bs = lam_b * rs[id];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[it][e] += bs * b[e];
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += x[it][e] * x[it][e];
    }
    float oscale = 1.f;
    if (out_norm) {
      ss = warp_sum(ss);
      oscale = rsqrtf(ss / (float)p.Do + p.eps);
    }
    T* orow = out + (size_t)pos * p.Do;
#pragma unroll
    for (int it = 0; it < CPL; ++it) {
      if (it * 32 + lane < p.n_chunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[it][e] *= oscale;
        Vec8<T>::stg(orow + (size_t)(it * 32 + lane) * kChunk, x[it]);
      }
    }
  };

  for (long long pos = gw; pos < p.N; pos += 2 * stride) {
    FwdLoad<T, CPL> a, b;
    const bool two = pos + stride < p.N;  // warp-uniform
    load(pos, a);
    if (two) load(pos + stride, b);
    finish(pos, a);
    if (two) finish(pos + stride, b);
  }
}

// ======================================================================================
// Backward
// ======================================================================================
// Finish one token row: given Du = sum over occurrences of d z[token part], the raw token row tv and
// its rms scale r_t, write d E_tok[v] and return this row's contribution to d lam_tok.
template <typename T, int CPL>
__device__ __forceinline__ float finish_tok_row(const EmbedParams& p, const ChunkMap (&cm)[CPL], int v,
                                                float (&Du)[CPL][8], const typename Vec8<T>::Raw (&traw)[CPL],
                                                float r_t, float lam_t) {
  const bool tok_norm = (p.flags & MOT_F_TOK_NORM) != 0;
  float dot = 0.f;  // <Du, tv>
#pragma unroll
  for (int it = 0; it < CPL; ++it) {
    if (cm[it].toff >= 0) {
      float tv[8];
      Vec8<T>::unpack(traw[it], tv);
#pragma unroll
      for (int e = 0; e < 8; ++e) dot += Du[it][e] * tv[e];
    }
  }
  const bool need_dot = tok_norm || (p.flags & MOT_F_HAS_LAMBDAS);
  if (need_dot) dot = warp_sum(dot);
  // d that = lam_t * Du ; dt = r_t * d that - tv * r_t^3 * mean(d that . tv)
  const float a = lam_t * r_t;
  const float b = tok_norm ? lam_t * r_t * r_t * r_t * dot / (float)p.Dt : 0.f;
  T* grow = reinterpret_cast<T*>(p.gE_tok) + (size_t)v * p.Dt;
#pragma unroll
  for (int it = 0; it < CPL; ++it) {
    if (cm[it].toff >= 0) {
      float tv[8], o[8];
      Vec8<T>::unpack(traw[it], tv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = a * Du[it][e] - b * tv[e];
      Vec8<T>::stg(grow + cm[it].toff, o);
    }
  }
  return r_t * dot;  // <Du, that>
}

template <typename T, int CPL, bool SMEM_ACC>
__global__ void __launch_bounds__(kBwdThreads, 1) mot_bwd_kernel(const EmbedParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  const size_t rs_bytes = align_up((size_t)p.Vb * sizeof(float), 128);
  const size_t tab_bytes = p.tab_smem ? align_up((size_t)p.Vb * p.bd * sizeof(T), 128) : 0;
  float* rs = reinterpret_cast<float*>(smem_raw);
  T* tab = reinterpret_cast<T*>(smem_raw + rs_bytes);
  float* acc = reinterpret_cast<float*>(smem_raw + rs_bytes + tab_bytes);
  const bool has_tok = p.combine != MOT_BYTES_ONLY;
  const bool has_bytes = p.combine != MOT_TOK_ONLY;
  const int nacc = p.Vb * p.acc_stride;
  if (has_bytes) {
    if (SMEM_ACC)
      for (int i = threadIdx.x; i < nacc; i += blockDim.x) acc[i] = 0.f;
    stage_byte_table<T>(p, tab, rs, &bar);  // ends with __syncthreads()
  }
  float* accp = SMEM_ACC ? acc : p.byte_acc;
  const int astride = SMEM_ACC ? p.acc_stride : p.bd;

  const int lane = lane_id();
  const int nw = blockDim.x >> 5;
  const int gw = blockIdx.x * nw + (threadIdx.x >> 5);
  const int W = gridDim.x * nw;

  ChunkMap cm[CPL];
#pragma unroll
  for (int it = 0; it < CPL; ++it) cm[it] = chunk_map(p, it * 32 + lane);

  const bool has_lam = (p.flags & MOT_F_HAS_LAMBDAS) != 0;
  float lam_t = 1.f, lam_b = 1.f;
  if (has_lam) {
    lam_t = __ldg(p.lam);
    lam_b = __ldg(p.lam + 1);
  }
  const float inv_pool = (p.combine == MOT_MEAN) ? 1.f / (float)p.bpt : 1.f;
  const float lam_b_eff = lam_b * inv_pool;
  const T* E_tok = reinterpret_cast<const T*>(p.E_tok);
  const T* gout = reinterpret_cast<const T*>(p.gout);
  const bool tok_norm = (p.flags & MOT_F_TOK_NORM) != 0;
  const bool out_norm = (p.flags & MOT_F_OUT_NORM) != 0;
  float dlam_t = 0.f, dlam_b = 0.f;  // per-lane partials

  using Raw = typename Vec8<T>::Raw;
  struct Occ {
    Raw graw[CPL];
    int idreg;
  };
  auto load_occ = [&](long long pos, Occ& o) {
    o.idreg = (has_bytes && lane < p.bpt) ? fetch_id(p, pos, lane) : 0;
    const T* grow = gout + (size_t)pos * p.Do;
#pragma unroll
    for (int it = 0; it < CPL; ++it)
      if (it * 32 + lane < p.n_chunks) o.graw[it] = Vec8<T>::ldg_raw(grow + (size_t)(it * 32 + lane) * kChunk);
  };

  // z chunk of this occurrence: token part (tscale * tv) + byte part (lam_b * rs * row)
  auto z_chunk = [&](int it, const Raw (&traw)[CPL], float tscale, int idreg, float (&z)[8], float (&bhat)[8]) {
    if (cm[it].toff >= 0) {
      Vec8<T>::unpack(traw[it], z);
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] *= tscale;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) bhat[e] = 0.f;
    if (has_bytes) {
      if (p.combine == MOT_MEAN) {
        for (int k = 0; k < p.bpt; ++k) {
          const int id = __shfl_sync(0xffffffffu, idreg, k);
          if (cm[it].slot == -2) {
            float b[8];
            Vec8<T>::unpack(tab_load<T>(p, tab, (size_t)id * p.bd + cm[it].boff), b);
            const float r = rs[id];
#pragma unroll
            for (int e = 0; e < 8; ++e) bhat[e] += r * b[e];
          }
        }
      } else {
        const int id = __shfl_sync(0xffffffffu, idreg, cm[it].slot & 31);
        if (cm[it].slot >= 0) {
          float b[8];
          Vec8<T>::unpack(tab_load<T>(p, tab, (size_t)id * p.bd + cm[it].boff), b);
          const float r = rs[id];
#pragma unroll
          for (int e = 0; e < 8; ++e) bhat[e] = r * b[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] += lam_b_eff * bhat[e];
    }
  };

  // one occurrence: accumulates d z[token part] into Du and lam_b * d z[byte part] into the byte accumulators
  auto process_occ = [&](const Occ& o, const Raw (&traw)[CPL], float tscale, float (&Du)[CPL][8]) {
    float g[CPL][8];
    float ss = 0.f, gz = 0.f;
#pragma unroll
    for (int it = 0; it < CPL; ++it) {
      if (it * 32 + lane < p.n_chunks) {
        Vec8<T>::unpack(o.graw[it], g[it]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) g[it][e] = 0.f;
      }
      if (out_norm) {
        float z[8], bhat[8];
        z_chunk(it, traw, tscale, o.idreg, z, bhat);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          ss += z[e] * z[e];
          gz += g[it][e] * z[e];
        }
      }
    }
    float r_o = 1.f, coef = 0.f;
    if (out_norm) {
      warp_sum2(ss, gz);
      r_o = rsqrtf(ss / (float)p.Do + p.eps);
      coef = r_o * r_o * r_o * gz / (float)p.Do;
    }
#pragma unroll
    for (int it = 0; it < CPL; ++it) {
      float z[8], bhat[8];
      z_chunk(it, traw, tscale, o.idreg, z, bhat);  // second pass: recompute instead of holding z in registers
      float dz[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) dz[e] = r_o * g[it][e] - coef * z[e];
      if (cm[it].toff >= 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) Du[it][e] += dz[e];
      }
      if (has_bytes) {
        if (has_lam) {
#pragma unroll
          for (int e = 0; e < 8; ++e) dlam_b += inv_pool * dz[e] * bhat[e];
        }
        float bv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) bv[e] = lam_b_eff * dz[e];
        if (p.combine == MOT_MEAN) {
          for (int k = 0; k < p.bpt; ++k) {
            const int id = __shfl_sync(0xffffffffu, o.idreg, k);
            if (cm[it].slot == -2) {
              float* a = accp + (size_t)id * astride + cm[it].boff;
              if (SMEM_ACC) smem_add8(a, bv); else gmem_add8(a, bv);
            }
          }
        } else {
          const int id = __shfl_sync(0xffffffffu, o.idreg, cm[it].slot & 31);
          if (cm[it].slot >= 0) {
            float* a = accp + (size_t)id * astride + cm[it].boff;
            if (SMEM_ACC) smem_add8(a, bv); else gmem_add8(a, bv);
          }
        }
      }
    }
  };

  if (has_tok) {
    // ---- phase Z: rows nobody gathered get zeros (the dense-grad contract of the reference) ----
    {
      T* G = reinterpret_cast<T*>(p.gE_tok);
      float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int vb = gw * 32; vb < p.V; vb += W * 32) {
        const int v = vb + lane;
        const bool empty = v < p.V && (p.off[v + 1] - p.off[v]) == 0;
        unsigned m = __ballot_sync(0xffffffffu, empty);
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          T* row = G + (size_t)(vb + j) * p.Dt;
          for (int c = lane; c < p.Dt / kChunk; c += 32) Vec8<T>::stg(row + c * kChunk, zero);
        }
      }
    }
    // ---- phase I: work items = (token row, <= L occurrences) ----
    const int n_items = p.item_off[p.V];
    for (int j = gw; j < n_items; j += W) {
      const int4 item = p.items[j];
      const int v = item.x, start = item.y, cnt = item.z, pslot = item.w;
      Raw traw[CPL];
      const T* trow = E_tok + (size_t)v * p.Dt;
      float ss_t = 0.f;
#pragma unroll
      for (int it = 0; it < CPL; ++it) {
        if (cm[it].toff >= 0) {
          traw[it] = Vec8<T>::ldg_raw(trow + cm[it].toff);
        } else {
          traw[it] = Vec8<T>::zero_raw();
        }
      }
      float r_t = 1.f;
      if (tok_norm) {
#pragma unroll
        for (int it = 0; it < CPL; ++it) {
          float tv[8];
          Vec8<T>::unpack(traw[it], tv);
#pragma unroll
          for (int e = 0; e < 8; ++e) ss_t += tv[e] * tv[e];
        }
        ss_t = warp_sum(ss_t);
        r_t = rsqrtf(ss_t / (float)p.Dt + p.eps);
      }
      const float tscale = lam_t * r_t;
      float Du[CPL][8];
#pragma unroll
      for (int it = 0; it < CPL; ++it)
#pragma unroll
        for (int e = 0; e < 8; ++e) Du[it][e] = 0.f;

      for (int k0 = 0; k0 < cnt; k0 += 32) {
        const int nb = min(32, cnt - k0);
        const int mypos = (lane < nb) ? p.order[start + k0 + lane] : 0;
        Occ cur, nxt;
        load_occ(__shfl_sync(0xffffffffu, mypos, 0), cur);
        for (int k = 0; k < nb; ++k) {
          const bool more = k + 1 < nb;  // warp-uniform
          if (more) load_occ(__shfl_sync(0xffffffffu, mypos, k + 1), nxt);
          process_occ(cur, traw, tscale, Du);
          if (more) cur = nxt;
        }
      }
      if (pslot < 0) {
        dlam_t += finish_tok_row<T, CPL>(p, cm, v, Du, traw, r_t, lam_t) * (lane == 0 ? 1.f : 0.f);
      } else {
        float* prow = p.partial + (size_t)pslot * p.Dt;
#pragma unroll
        for (int it = 0; it < CPL; ++it) {
          if (cm[it].toff >= 0) {
            *reinterpret_cast<float4*>(prow + cm[it].toff) = make_float4(Du[it][0], Du[it][1], Du[it][2], Du[it][3]);
            *reinterpret_cast<float4*>(prow + cm[it].toff + 4) = make_float4(Du[it][4], Du[it][5], Du[it][6], Du[it][7]);
          }
        }
      }
    }
  } else {
    // bytes-only: no token table, plain position loop
    Raw traw[CPL];
#pragma unroll
    for (int it = 0; it < CPL; ++it) traw[it] = Vec8<T>::zero_raw();
    float Du[CPL][8];
#pragma unroll
    for (int it = 0; it < CPL; ++it)
#pragma unroll
      for (int e = 0; e < 8; ++e) Du[it][e] = 0.f;
    for (long long pos = gw; pos < p.N; pos += W) {
      Occ cur;
      load_occ(pos, cur);
      process_occ(cur, traw, 0.f, Du);
    }
  }

  // ---- CTA epilogue: flush the byte accumulators and the lambda partials ----
  if (has_bytes && SMEM_ACC) {
    __syncthreads();
    const int c4 = p.bd / 4;  // float4 groups per row (bd is a multiple of 8)
    for (int i = threadIdx.x; i < p.Vb * c4; i += blockDim.x) {
      const int r = i / c4, c = (i - r * c4) * 4;
      const float2 lo = *reinterpret_cast<const float2*>(acc + (size_t)r * p.acc_stride + c);
      const float2 hi = *reinterpret_cast<const float2*>(acc + (size_t)r * p.acc_stride + c + 2);
      if (lo.x != 0.f || lo.y != 0.f || hi.x != 0.f || hi.y != 0.f)
        atomicAdd(reinterpret_cast<float4*>(p.byte_acc + (size_t)r * p.bd + c), make_float4(lo.x, lo.y, hi.x, hi.y));
    }
  }
  if (has_lam) {
    dlam_t = warp_sum(dlam_t);
    dlam_b = warp_sum(dlam_b);
    if (lane == 0) {
      atomicAdd(p.lam_acc, dlam_t);
      atomicAdd(p.lam_acc + 1, dlam_b);
    }
  }
}

// Finalize: (a) hot token rows: sum their fp32 partials in order, apply the token-norm backward, write
// the row; (b) byte table: apply the byte-norm backward to the accumulated rows and cast; (c) lambdas.
template <typename T>
__global__ void __launch_bounds__(256) mot_bwd_finalize_kernel(const EmbedParams p) {
  const int lane = lane_id();
  const int nw = blockDim.x >> 5;
  const int gw = blockIdx.x * nw + (threadIdx.x >> 5);
  const int W = gridDim.x * nw;
  const bool has_tok = p.combine != MOT_BYTES_ONLY;
  const bool has_bytes = p.combine != MOT_TOK_ONLY;
  const bool has_lam = (p.flags & MOT_F_HAS_LAMBDAS) != 0;
  float lam_t = 1.f;
  if (has_lam) lam_t = __ldg(p.lam);
  float dlam_t = 0.f;
  if (has_tok) {
    const bool tok_norm = (p.flags & MOT_F_TOK_NORM) != 0;
    const int n_hot = p.hot_off[p.V];
    const T* E_tok = reinterpret_cast<const T*>(p.E_tok);
    T* G = reinterpret_cast<T*>(p.gE_tok);
    for (int h = gw; h < n_hot; h += W) {
      const int4 hr = p.hot_rows[h];
      const int v = hr.x, ps = hr.y, nch = hr.z;
      const T* trow = E_tok + (size_t)v * p.Dt;
      // pass 1: dot = <Du, tv>, ss = |tv|^2
      float dot = 0.f, ss = 0.f;
      for (int c = lane; c < p.Dt / kChunk; c += 32) {
        float tv[8], du[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        Vec8<T>::unpack(Vec8<T>::ldg_raw(trow + c * kChunk), tv);
        for (int j = 0; j < nch; ++j) {
          const float* pr = p.partial + (size_t)(ps + j) * p.Dt + c * kChunk;
          const float4 a = *reinterpret_cast<const float4*>(pr), b = *reinterpret_cast<const float4*>(pr + 4);
          du[0] += a.x; du[1] += a.y; du[2] += a.z; du[3] += a.w;
          du[4] += b.x; du[5] += b.y; du[6] += b.z; du[7] += b.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dot += du[e] * tv[e];
          ss += tv[e] * tv[e];
        }
      }
      warp_sum2(dot, ss);
      const float r_t = tok_norm ? rsqrtf(ss / (float)p.Dt + p.eps) : 1.f;
      const float a_ = lam_t * r_t;
      const float b_ = tok_norm ? lam_t * r_t * r_t * r_t * dot / (float)p.Dt : 0.f;
      for (int c = lane; c < p.Dt / kChunk; c += 32) {
        float tv[8], du[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, o[8];
        Vec8<T>::unpack(Vec8<T>::ldg_raw(trow + c * kChunk), tv);
        for (int j = 0; j < nch; ++j) {
          const float* pr = p.partial + (size_t)(ps + j) * p.Dt + c * kChunk;
          const float4 a = *reinterpret_cast<const float4*>(pr), b = *reinterpret_cast<const float4*>(pr + 4);
          du[0] += a.x; du[1] += a.y; du[2] += a.z; du[3] += a.w;
          du[4] += b.x; du[5] += b.y; du[6] += b.z; du[7] += b.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = a_ * du[e] - b_ * tv[e];
        Vec8<T>::stg(G + (size_t)v * p.Dt + c * kChunk, o);
      }
      if (lane == 0) dlam_t += r_t * dot;
    }
  }
  if (has_bytes) {
    const bool bn = (p.flags & MOT_F_BYTE_NORM) != 0;
    const T* E_byte = reinterpret_cast<const T*>(p.E_byte);
    T* G = reinterpret_cast<T*>(p.gE_byte);
    for (int r = gw; r < p.Vb; r += W) {
      float dot = 0.f, ss = 0.f;
      if (bn) {
        for (int c = lane; c < p.bd / kChunk; c += 32) {
          float ev[8];
          Vec8<T>::unpack(Vec8<T>::ldg_raw(E_byte + (size_t)r * p.bd + c * kChunk), ev);
          const float* a = p.byte_acc + (size_t)r * p.bd + c * kChunk;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            dot += a[e] * ev[e];
            ss += ev[e] * ev[e];
          }
        }
        warp_sum2(dot, ss);
      }
      const float rsr = bn ? rsqrtf(ss / (float)p.bd + p.eps) : 1.f;
      const float b_ = bn ? rsr * rsr * rsr * dot / (float)p.bd : 0.f;
      for (int c = lane; c < p.bd / kChunk; c += 32) {
        float ev[8], o[8];
        Vec8<T>::unpack(Vec8<T>::ldg_raw(E_byte + (size_t)r * p.bd + c * kChunk), ev);
        const float* a = p.byte_acc + (size_t)r * p.bd + c * kChunk;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rsr * a[e] - b_ * ev[e];
        Vec8<T>::stg(G + (size_t)r * p.bd + c * kChunk, o);
      }
    }
  }
  if (has_lam) {
    dlam_t = warp_sum(dlam_t);
    if (lane == 0 && dlam_t != 0.f) atomicAdd(p.lam_acc, dlam_t);
  }
}


inline size_t smem_bytes(const EmbedParams& p, size_t esz, bool tab, bool acc) {
  if (p.combine == MOT_TOK_ONLY) return 0;
  return align_up((size_t)p.Vb * 4, 128) + (tab ? align_up((size_t)p.Vb * p.bd * esz, 128) : 0) +
         (acc ? align_up((size_t)p.Vb * (p.bd + 2) * 4, 128) : 0);
}

template <typename T, int CPL>
static int launch_fwd(const EmbedParams& p_in, cudaStream_t s) {
  int sms = 0, optin = 0;
  if (int rc = device_props(&sms, &optin)) return rc;
  EmbedParams p = p_in;
  p.tab_smem = 1;
  size_t smem = smem_bytes(p, sizeof(T), true, false);
  if (smem + 1024 > (size_t)optin) {  // table larger than one SM's shared memory: gather rows through L1/L2
    p.tab_smem = 0;
    smem = smem_bytes(p, sizeof(T), false, false);
  }
  auto kern = mot_fwd_kernel<T, CPL>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return check_launch();
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFwdThreads, smem);
  if (occ < 1) occ = 1;
  const long long warps_needed = (p.N + 1) / 2;  // two positions per warp iteration
  long long blocks = (warps_needed + (kFwdThreads / 32) - 1) / (kFwdThreads / 32);
  const long long cap = (long long)sms * occ;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (g_prof_fwd_start) cudaEventRecord(g_prof_fwd_start, s);
  kern<<<(unsigned)blocks, kFwdThreads, smem, s>>>(p);
  if (g_prof_fwd_stop) cudaEventRecord(g_prof_fwd_stop, s);
  count_launch();
  return check_launch();
}


template <typename T, int CPL>
static int launch_bwd(const EmbedParams& p_in, cudaStream_t s) {
  int sms = 0, optin = 0;
  if (int rc = device_props(&sms, &optin)) return rc;
  EmbedParams p = p_in;
  // preference: table + fp32 accumulators in shared memory; else accumulators in the L2-resident
  // scratch (fp32 atomics); else the table through L1/L2 as well
  bool smem_acc = true;
  p.tab_smem = 1;
  p.acc_stride = p.bd + 2;
  static const char* acc_env = getenv("MOT_BWD_ACC");  // debug knob: "global" forces the L2-atomics path
  const bool force_global = acc_env && acc_env[0] == 'g';
  size_t smem = smem_bytes(p, sizeof(T), true, true);
  if (force_global || smem + 1024 > (size_t)optin) {
    smem_acc = false;
    smem = smem_bytes(p, sizeof(T), true, false);
    if (smem + 1024 > (size_t)optin) {
      p.tab_smem = 0;
      smem = smem_bytes(p, sizeof(T), false, false);
    }
  }
  auto kern = smem_acc ? mot_bwd_kernel<T, CPL, true> : mot_bwd_kernel<T, CPL, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return check_launch();
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBwdThreads, smem);
  if (occ < 1) occ = 1;
  long long blocks = (long long)sms * occ;
  if (g_prof_start) cudaEventRecord(g_prof_start, s);
  kern<<<(unsigned)blocks, kBwdThreads, smem, s>>>(p);
  if (g_prof_stop) cudaEventRecord(g_prof_stop, s);
  count_launch();
  return check_launch();
}



// per-translation-unit dispatchers (one TU per element type so nvcc compiles them in parallel)
int dispatch_fwd_bf16(const EmbedParams& p, cudaStream_t s);
int dispatch_fwd_f32(const EmbedParams& p, cudaStream_t s);
int dispatch_bwd_bf16(const EmbedParams& p, cudaStream_t s);
int dispatch_bwd_f32(const EmbedParams& p, cudaStream_t s);
int launch_finalize_bf16(const EmbedParams& p, int blocks, cudaStream_t s);
int launch_finalize_f32(const EmbedParams& p, int blocks, cudaStream_t s);

}  // namespace mot
