// Small HBM-bound helpers of the projection variants that used to be eager PyTorch calls on the product path:
//   mot_cast_f32_bf16 : CastedLinear's per-call `self.weight.type_as(x)` (spt/train_gpt.py:185-186): the fp32 master weight
//                       of the mixin projection rounded to bf16 (round-to-nearest-even, like .to(bfloat16));
//   mot_colsum        : autograd of F.linear's bias (mathblations/model.py:261,268: nn.Linear with bias): the column sums
//                       of dY [n, dim] in fp32, in a fixed order (deterministic: no atomics).
#include "mot_common.cuh"

namespace mot {

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* in, long long n, __nv_bfloat16* out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long groups = n >> 3;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const uint4 a = ldg_nc_16(in + (g << 3)), b = ldg_nc_16(in + (g << 3) + 4);
    uint4 r;
    r.x = f32x2_to_bf16x2(__uint_as_float(a.x), __uint_as_float(a.y));
    r.y = f32x2_to_bf16x2(__uint_as_float(a.z), __uint_as_float(a.w));
    r.z = f32x2_to_bf16x2(__uint_as_float(b.x), __uint_as_float(b.y));
    r.w = f32x2_to_bf16x2(__uint_as_float(b.z), __uint_as_float(b.w));
    stg_16(out + (g << 3), r);
  }
  for (long long i = (groups << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16_rn(in[i]);
}

// One CTA owns 64 columns (a lane two adjacent ones: 128-byte / 256-byte warp rows), its 32 warps walk the rows with
// stride 32 and meet in shared memory in warp order: the summation order is fixed by the shape alone.
template <typename T>
__global__ void __launch_bounds__(1024) colsum_kernel(const T* x, long long n_rows, int dim, float* out) {
  __shared__ float part[32][64];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;
  float a0 = 0.f, a1 = 0.f;
  if (c < dim) {
    for (long long r = warp; r < n_rows; r += 32) {
      const T* p = x + r * dim + c;
      if (sizeof(T) == 2) {
        float lo, hi;
        bf16x2_to_f32(*reinterpret_cast<const uint32_t*>(p), lo, hi);
        a0 += lo; a1 += hi;
      } else {
        const float2 v = *reinterpret_cast<const float2*>(p);
        a0 += v.x; a1 += v.y;
      }
    }
  }
  part[warp][lane * 2] = a0;
  part[warp][lane * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + (int)threadIdx.x < dim) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) s += part[w][threadIdx.x];
    out[blockIdx.x * 64 + threadIdx.x] = s;
  }
}

}  // namespace mot

extern "C" int mot_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  if (n < 0) return MOT_ERR_BAD_ARG;
  if (n == 0) return MOT_OK;
  if (!in || !out) return MOT_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(in) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return MOT_ERR_MISALIGNED;
  int sms = 0, optin = 0;
  if (int rc = mot::device_props(&sms, &optin)) return rc;
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > sms * 8LL) blocks = sms * 8LL;
  if (blocks < 1) blocks = 1;
  mot::launch_pdl(mot::cast_f32_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), in,
                  (long long)n, reinterpret_cast<__nv_bfloat16*>(out));
  mot::count_launch();
  return mot::check_launch();
}

extern "C" int mot_colsum(const void* x, float* out, int64_t n_rows, int32_t dim, int32_t dtype, void* stream) {
  if (n_rows < 0 || dim <= 0) return MOT_ERR_BAD_ARG;
  if (dtype != MOT_BF16 && dtype != MOT_F32) return MOT_ERR_UNSUPPORTED;
  if (dim % 2) return MOT_ERR_MISALIGNED;
  if (!out || (n_rows > 0 && !x)) return MOT_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(x) & 7u) return MOT_ERR_MISALIGNED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const dim3 g((unsigned)((dim + 63) / 64)), b(1024);
  if (dtype == MOT_BF16)
    mot::launch_pdl(mot::colsum_kernel<__nv_bfloat16>, g, b, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), (long long)n_rows, (int)dim, out);
  else
    mot::launch_pdl(mot::colsum_kernel<float>, g, b, 0, s, reinterpret_cast<const float*>(x), (long long)n_rows, (int)dim, out);
  mot::count_launch();
  return mot::check_launch();
}
