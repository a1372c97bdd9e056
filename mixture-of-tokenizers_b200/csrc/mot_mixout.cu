// Output side of the path: the expand that turns token rows into byte rows (mirror image of the input mix).
//
// Replaces ByteMixoutCopy.forward's `einops.repeat(x, "... T D -> ... (T bpt) D")` (spt/train_gpt.py:493) and its
// autograd (sum of the bpt copies).  ByteMixoutSplit's `rearrange "... T (bpt D) -> ... (T bpt) D"` (:516) is a pure
// view of contiguous rows and needs no kernel.  HBM bound: forward reads N*D, writes N*bpt*D elements; backward the
// reverse, summing in fp32 and rounding once.  One thread owns one 16-byte piece of a token row.
#include "mot_common.cuh"

namespace mot {

template <typename T>
__global__ void __launch_bounds__(256) mixout_copy_fwd_kernel(const T* x, long long n_pieces, int pieces_per_row, int bpt,
                                                             T* y) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int E = 16 / sizeof(T);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pieces; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / pieces_per_row;
    const int c = (int)(i - row * pieces_per_row);
    const uint4 v = ldg_nc_16(x + i * E);
    T* o = y + (row * bpt * (long long)pieces_per_row + c) * E;
    for (int k = 0; k < bpt; ++k) stg_16(o + (long long)k * pieces_per_row * E, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) mixout_copy_bwd_kernel(const T* gy, long long n_pieces, int pieces_per_row, int bpt,
                                                             T* gx) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int E = 16 / sizeof(T);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pieces; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / pieces_per_row;
    const int c = (int)(i - row * pieces_per_row);
    const T* s = gy + (row * bpt * (long long)pieces_per_row + c) * E;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (; k + 4 <= bpt; k += 4) {  // four independent 16-byte loads in flight
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ldg_nc_16(s + (long long)(k + j) * pieces_per_row * E);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (sizeof(T) == 2) {
          float f[8];
          bf16x2_to_f32(v[j].x, f[0], f[1]); bf16x2_to_f32(v[j].y, f[2], f[3]);
          bf16x2_to_f32(v[j].z, f[4], f[5]); bf16x2_to_f32(v[j].w, f[6], f[7]);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += f[e];
        } else {
          acc[0] += __uint_as_float(v[j].x); acc[1] += __uint_as_float(v[j].y);
          acc[2] += __uint_as_float(v[j].z); acc[3] += __uint_as_float(v[j].w);
        }
      }
    }
    for (; k < bpt; ++k) {
      const uint4 v = ldg_nc_16(s + (long long)k * pieces_per_row * E);
      if (sizeof(T) == 2) {
        float f[8];
        bf16x2_to_f32(v.x, f[0], f[1]); bf16x2_to_f32(v.y, f[2], f[3]);
        bf16x2_to_f32(v.z, f[4], f[5]); bf16x2_to_f32(v.w, f[6], f[7]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      } else {
        acc[0] += __uint_as_float(v.x); acc[1] += __uint_as_float(v.y);
        acc[2] += __uint_as_float(v.z); acc[3] += __uint_as_float(v.w);
      }
    }
    uint4 r;
    if (sizeof(T) == 2) {
      r.x = f32x2_to_bf16x2(acc[0], acc[1]); r.y = f32x2_to_bf16x2(acc[2], acc[3]);
      r.z = f32x2_to_bf16x2(acc[4], acc[5]); r.w = f32x2_to_bf16x2(acc[6], acc[7]);
    } else {
      r.x = __float_as_uint(acc[0]); r.y = __float_as_uint(acc[1]); r.z = __float_as_uint(acc[2]); r.w = __float_as_uint(acc[3]);
    }
    stg_16(gx + i * E, r);
  }
}

template <typename T>
static int launch_mixout(const void* in, void* out, long long n_rows, int dim, int bpt, bool backward, cudaStream_t s) {
  int sms = 0, optin = 0;
  if (int rc = device_props(&sms, &optin)) return rc;
  constexpr int E = 16 / sizeof(T);
  const int ppr = dim / E;
  const long long n_pieces = n_rows * ppr;
  long long blocks = (n_pieces + 255) / 256;
  if (blocks > sms * 16LL) blocks = sms * 16LL;
  if (backward) launch_pdl(mixout_copy_bwd_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, s, (const T*)in, n_pieces, ppr, bpt, (T*)out);
  else launch_pdl(mixout_copy_fwd_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, s, (const T*)in, n_pieces, ppr, bpt, (T*)out);
  count_launch();
  return check_launch();
}

}  // namespace mot

static int mixout_args(const void* a, const void* b, int64_t n_rows, int32_t dim, int32_t bpt, int32_t dtype) {
  if (n_rows < 0 || dim <= 0 || bpt <= 0) return MOT_ERR_BAD_ARG;
  if (dtype != MOT_BF16 && dtype != MOT_F32) return MOT_ERR_UNSUPPORTED;
  if (dim % (dtype == MOT_BF16 ? 8 : 4)) return MOT_ERR_MISALIGNED;
  if (n_rows == 0) return MOT_OK;
  if (!a || !b) return MOT_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(a) & 15u) || (reinterpret_cast<uintptr_t>(b) & 15u)) return MOT_ERR_MISALIGNED;
  return MOT_OK;
}

extern "C" int mot_mixout_copy_fwd(const void* x, void* y, int64_t n_rows, int32_t dim, int32_t bpt, int32_t dtype, void* stream) {
  if (int rc = mixout_args(x, y, n_rows, dim, bpt, dtype)) return rc;
  if (n_rows == 0) return MOT_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return dtype == MOT_BF16 ? mot::launch_mixout<__nv_bfloat16>(x, y, n_rows, dim, bpt, false, s)
                           : mot::launch_mixout<float>(x, y, n_rows, dim, bpt, false, s);
}

extern "C" int mot_mixout_copy_bwd(const void* grad_y, void* grad_x, int64_t n_rows, int32_t dim, int32_t bpt, int32_t dtype,
                                   void* stream) {
  if (int rc = mixout_args(grad_y, grad_x, n_rows, dim, bpt, dtype)) return rc;
  if (n_rows == 0) return MOT_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return dtype == MOT_BF16 ? mot::launch_mixout<__nv_bfloat16>(grad_y, grad_x, n_rows, dim, bpt, true, s)
                           : mot::launch_mixout<float>(grad_y, grad_x, n_rows, dim, bpt, true, s);
}
