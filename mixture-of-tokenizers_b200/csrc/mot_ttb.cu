// Integer half of the path: token ids -> padded byte ids.
//
// Replaces tokens_to_bytes (spt/data_creation.py:61-67; runs/7:444-450): `emb(tokens).to(int64)`
// on a float nn.Embedding that holds the ttb table.  Here the table is int16 [V, bpt] (or the
// reference's fp32 / bf16 containers, read back with the same truncation), one thread moves one
// 8-byte group of four ids, so a warp reads 256 B of table rows and writes 512 B / 1 KB of output
// fully coalesced.
#include "mot_common.cuh"

namespace mot {

template <int TTB, typename OutT>
__global__ void __launch_bounds__(256) ttb_expand_kernel(const int32_t* tok, long long n, const void* ttb,
                                                        int V, int bpt, OutT* out) {
  const long long total = n * bpt;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (TTB == MOT_TTB_I16 && (bpt & 3) == 0) {
    // four ids per thread: one 8-byte table load, one 16/32-byte store
    const long long groups = total >> 2;
    const int gpt = bpt >> 2;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
      const long long t = g / gpt;
      const int k = (int)(g - t * gpt);
      const int tv = min(max(ld_g(tok + t), 0), V - 1);
      const uint2 r = ld_g(reinterpret_cast<const uint2*>(reinterpret_cast<const short*>(ttb) + (size_t)tv * bpt) + k);
      const OutT a = (OutT)(short)(r.x & 0xffffu), b = (OutT)(short)(r.x >> 16), c = (OutT)(short)(r.y & 0xffffu),
                 d = (OutT)(short)(r.y >> 16);
      OutT* o = out + (g << 2);
      o[0] = a; o[1] = b; o[2] = c; o[3] = d;
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long t = i / bpt;
    const int k = (int)(i - t * bpt);
    const int tv = min(max(ld_g(tok + t), 0), V - 1);
    const size_t e = (size_t)tv * bpt + k;
    long long id;
    if (TTB == MOT_TTB_I16) id = ld_g(reinterpret_cast<const short*>(ttb) + e);
    else if (TTB == MOT_TTB_F32) id = (long long)ld_g(reinterpret_cast<const float*>(ttb) + e);  // trunc, like .to(int64)
    else id = (long long)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ttb)[e]);
    out[i] = (OutT)id;
  }
}

template <int TTB>
static int launch_expand(const int32_t* tok, long long n, const void* ttb, int V, int bpt, void* out, int out_i64,
                         cudaStream_t s) {
  int sms = 0, optin = 0;
  if (int rc = device_props(&sms, &optin)) return rc;
  const long long total = n * bpt;
  long long blocks = (total / 4 + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  if (blocks < 1) blocks = 1;
  if (out_i64) ttb_expand_kernel<TTB, long long><<<(unsigned)blocks, 256, 0, s>>>(tok, n, ttb, V, bpt, reinterpret_cast<long long*>(out));
  else ttb_expand_kernel<TTB, int><<<(unsigned)blocks, 256, 0, s>>>(tok, n, ttb, V, bpt, reinterpret_cast<int*>(out));
  count_launch();
  return check_launch();
}

// mathblations analogue of the ttb expansion, computed arithmetically (mathblations/data.py:92-109): a number token
// becomes its decimal digits right-aligned in dpt slots padded with 13; the operator / equals / pad tokens become
// 10 / 11 / 12 in the last slot.
template <typename TokT, typename OutT>
__global__ void __launch_bounds__(256) tokens_to_digits_kernel(const TokT* tok, long long n, int dpt, long long op_token,
                                                              long long eq_token, long long pad_token, OutT* out) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long t = (long long)tok[i];
    OutT* o = out + i * dpt;
    const int special = t == op_token ? 10 : (t == eq_token ? 11 : (t == pad_token ? 12 : -1));
    for (int j = dpt - 1; j >= 0; --j) {
      int d = 13;
      if (special >= 0) {
        if (j == dpt - 1) d = special;
      } else if (j == dpt - 1 || t > 0) {  // "0" is one digit; leading positions stay pad
        d = (int)(t % 10);
        t /= 10;
      }
      o[j] = (OutT)d;
    }
  }
}

// Shard tokens are stored as uint16 (spt/train_gpt.py:628-638; modded-nanogpt/data/fineweb.py): widen them to the int32
// ids every kernel of the path takes, on the device, so that the host -> device copy moves 2 bytes per token instead of
// the 4 of the reference's host-side `.to(torch.int32)` (spt/train_gpt.py:646).  8 tokens per thread: one 16-byte load,
// two 16-byte stores; HBM bound at 6 bytes per token.
__global__ void __launch_bounds__(256) tokens_widen_u16_kernel(const unsigned short* in, long long n, int* out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long groups = n >> 3;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const uint4 r = ldg_nc_16(in + (g << 3));
    uint4 a, b;
    a.x = r.x & 0xffffu; a.y = r.x >> 16; a.z = r.y & 0xffffu; a.w = r.y >> 16;
    b.x = r.z & 0xffffu; b.y = r.z >> 16; b.z = r.w & 0xffffu; b.w = r.w >> 16;
    stg_16(out + (g << 3), a);
    stg_16(out + (g << 3) + 4, b);
  }
  for (long long i = (groups << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int)in[i];
}


// ttb table build (modded-nanogpt/create_ttb.py:18-31) as an integer gather: row v of the table is the first
// min(bpt, len_v) character ids of token v's string, padded on the left or on the right with `pad_byte`; a token whose
// string is the end-of-text marker becomes bpt copies of `eot_byte` (:20-22).  The strings arrive as one flat id
// stream `chars` with offsets `offs` [n_rows + 1] (a negative length marks an end-of-text row).  One thread per table
// entry: out[v, k] is either a pad or ONE gathered id, so all stores are coalesced 2-byte writes of consecutive k.
__global__ void __launch_bounds__(256) ttb_build_kernel(const short* chars, const int* offs,
                                                       const unsigned char* is_eot, int n_rows, int bpt, int pad_left,
                                                       int pad_byte, int eot_byte, short* out) {
  const long long total = (long long)n_rows * bpt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / bpt), k = (int)(i - (long long)v * bpt);
    int id = pad_byte;
    if (is_eot != nullptr && is_eot[v]) {
      id = eot_byte;
    } else {
      const int a = ld_g(offs + v), len = min(ld_g(offs + v + 1) - a, bpt);  // keeps the FIRST bpt characters
      const int j = pad_left ? k - (bpt - len) : k;                            // index into the kept characters
      if (j >= 0 && j < len) id = ld_g(chars + a + j);
    }
    out[i] = (short)id;
  }
}

// Re-pad / truncate / flip the pad side of an existing table (the scaled-pre-train tables ttb_{16,18,20}_{left,right}_pad
// are the same strings under another (bpt, pad_position), spt/train_gpt.py:665-672).  A row's characters are its non-pad
// entries in order; rows that are all `eot_byte` stay all `eot_byte`.  One warp per row (bpt <= 32): ballot of the
// non-pad lanes gives every character its rank, i.e. its place in the re-padded row.
__global__ void __launch_bounds__(256) ttb_repad_kernel(const short* in, int n_rows, int bpt_in, int bpt_out, int pad_left,
                                                       int pad_byte, int eot_byte, short* out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < n_rows; v += warps) {
    const int id = lane < bpt_in ? (int)ld_g(in + (size_t)v * bpt_in + lane) : pad_byte;
    const unsigned live = __ballot_sync(0xffffffffu, lane < bpt_in && id != pad_byte);
    const unsigned eot = __ballot_sync(0xffffffffu, lane >= bpt_in || id == eot_byte);
    short* o = out + (size_t)v * bpt_out;
    if (eot == 0xffffffffu) {
      if (lane < bpt_out) o[lane] = (short)eot_byte;
      continue;
    }
    const int n = min(__popc(live), bpt_out);           // characters kept (the first n)
    if (lane < bpt_out) o[lane] = (short)pad_byte;
    __syncwarp();
    const int rank = __popc(live & ((1u << lane) - 1u));
    if (((live >> lane) & 1u) && rank < n) o[pad_left ? bpt_out - n + rank : rank] = (short)id;
  }
}

}  // namespace mot

extern "C" int mot_tokens_widen_u16(const void* tok_u16, int64_t n, int32_t* out, void* stream) {
  if (n < 0) return MOT_ERR_BAD_ARG;
  if (n == 0) return MOT_OK;
  if (!tok_u16 || !out) return MOT_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(tok_u16) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return MOT_ERR_MISALIGNED;
  int sms = 0, optin = 0;
  if (int rc = mot::device_props(&sms, &optin)) return rc;
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > sms * 8LL) blocks = sms * 8LL;
  if (blocks < 1) blocks = 1;
  mot::launch_pdl(mot::tokens_widen_u16_kernel, dim3((unsigned)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                  reinterpret_cast<const unsigned short*>(tok_u16), (long long)n, out);
  mot::count_launch();
  return mot::check_launch();
}

extern "C" int mot_tokens_to_digits(const void* tok, int64_t n, int32_t tok_i64, int32_t dpt, int64_t op_token, int64_t eq_token,
                                    int64_t pad_token, void* out, int32_t out_i64, void* stream) {
  if (n < 0 || dpt <= 0) return MOT_ERR_BAD_ARG;
  if (n == 0) return MOT_OK;
  if (!tok || !out) return MOT_ERR_BAD_ARG;
  int sms = 0, optin = 0;
  if (int rc = mot::device_props(&sms, &optin)) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  long long blocks = (n + 255) / 256;
  if (blocks > sms * 8LL) blocks = sms * 8LL;
  const dim3 g((unsigned)blocks), b(256);
  using namespace mot;
  if (tok_i64 && out_i64) launch_pdl(tokens_to_digits_kernel<long long, long long>, g, b, 0, s, (const long long*)tok, (long long)n, (int)dpt, (long long)op_token, (long long)eq_token, (long long)pad_token, (long long*)out);
  else if (tok_i64) launch_pdl(tokens_to_digits_kernel<long long, int>, g, b, 0, s, (const long long*)tok, (long long)n, (int)dpt, (long long)op_token, (long long)eq_token, (long long)pad_token, (int*)out);
  else if (out_i64) launch_pdl(tokens_to_digits_kernel<int, long long>, g, b, 0, s, (const int*)tok, (long long)n, (int)dpt, (long long)op_token, (long long)eq_token, (long long)pad_token, (long long*)out);
  else launch_pdl(tokens_to_digits_kernel<int, int>, g, b, 0, s, (const int*)tok, (long long)n, (int)dpt, (long long)op_token, (long long)eq_token, (long long)pad_token, (int*)out);
  count_launch();
  return check_launch();
}

extern "C" int mot_ttb_expand(const int32_t* tok, int64_t n, const void* ttb, int32_t tok_vocab, int32_t bpt,
                              int32_t ttb_dtype, void* out, int32_t out_i64, void* stream) {
  if (n < 0 || tok_vocab <= 0 || bpt <= 0) return MOT_ERR_BAD_ARG;
  if (n == 0) return MOT_OK;
  if (!tok || !ttb || !out) return MOT_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(ttb) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return MOT_ERR_MISALIGNED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (ttb_dtype) {
    case MOT_TTB_I16: return mot::launch_expand<MOT_TTB_I16>(tok, n, ttb, tok_vocab, bpt, out, out_i64, s);
    case MOT_TTB_F32: return mot::launch_expand<MOT_TTB_F32>(tok, n, ttb, tok_vocab, bpt, out, out_i64, s);
    case MOT_TTB_BF16: return mot::launch_expand<MOT_TTB_BF16>(tok, n, ttb, tok_vocab, bpt, out, out_i64, s);
  }
  return MOT_ERR_UNSUPPORTED;
}

extern "C" int mot_ttb_build(const int16_t* chars, const int32_t* offs, const uint8_t* is_eot, int32_t n_rows, int32_t bpt,
                             int32_t pad_left, int32_t pad_byte, int32_t eot_byte, int16_t* out, void* stream) {
  if (n_rows < 0 || bpt <= 0) return MOT_ERR_BAD_ARG;
  if (n_rows == 0) return MOT_OK;
  if (!offs || !out) return MOT_ERR_BAD_ARG;   // chars may be null when every string is empty
  int sms = 0, optin = 0;
  if (int rc = mot::device_props(&sms, &optin)) return rc;
  long long blocks = ((long long)n_rows * bpt + 255) / 256;
  if (blocks > sms * 8LL) blocks = sms * 8LL;
  mot::ttb_build_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      chars, offs, is_eot, n_rows, bpt, pad_left, pad_byte, eot_byte, out);
  mot::count_launch();
  return mot::check_launch();
}

extern "C" int mot_ttb_repad(const int16_t* table, int32_t n_rows, int32_t bpt_in, int32_t bpt_out, int32_t pad_left,
                             int32_t pad_byte, int32_t eot_byte, int16_t* out, void* stream) {
  if (n_rows < 0 || bpt_in <= 0 || bpt_out <= 0) return MOT_ERR_BAD_ARG;
  if (bpt_in > 32 || bpt_out > 32) return MOT_ERR_UNSUPPORTED;
  if (n_rows == 0) return MOT_OK;
  if (!table || !out || table == out) return MOT_ERR_BAD_ARG;
  int sms = 0, optin = 0;
  if (int rc = mot::device_props(&sms, &optin)) return rc;
  long long blocks = ((long long)n_rows + 7) / 8;
  if (blocks > sms * 8LL) blocks = sms * 8LL;
  mot::ttb_repad_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, n_rows, bpt_in, bpt_out, pad_left, pad_byte, eot_byte, out);
  mot::count_launch();
  return mot::check_launch();
}
