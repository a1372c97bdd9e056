// Library plumbing: error strings, launch accounting, device properties.
#include <atomic>
#include <mutex>
#include <csignal>
#include <cstdio>
#include <unistd.h>

#include "mot_common.cuh"

namespace mot {

thread_local cudaError_t g_last_cuda_error = cudaSuccess;
static std::atomic<long long> g_launches{0};
long long* g_trace = nullptr;
cudaEvent_t g_prof_fwd_start = nullptr, g_prof_fwd_stop = nullptr, g_prof_start = nullptr, g_prof_stop = nullptr;

#ifdef MOT_CHECK
static long long* g_chk_host = nullptr;
static void chk_abort_handler(int) {
  if (g_chk_host && g_chk_host[0]) {
    char buf[256];
    int n = snprintf(buf, sizeof buf, "\nMOT_CHECK violation: source line %lld values (%lld, %lld) block %lld thread %lld v_lo<<32|v_hi %lld plan_early %lld R %lld grid %lld\n",
                     g_chk_host[1], g_chk_host[2], g_chk_host[3], g_chk_host[4], g_chk_host[5], g_chk_host[6], g_chk_host[7] >> 40,
                     (g_chk_host[7] >> 20) & 0xfffff, g_chk_host[7] & 0xfffff);
    if (n > 0) (void)!write(2, buf, (size_t)n);
  }
  if (g_chk_host) g_chk_host[0] = 0;
  signal(SIGABRT, SIG_DFL);
  raise(SIGABRT);
}
long long* chk_record() {
  static long long* dev = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (cudaHostAlloc(reinterpret_cast<void**>(&g_chk_host), 64, cudaHostAllocMapped) == cudaSuccess) {
      for (int i = 0; i < 8; ++i) g_chk_host[i] = 0;
      cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev), g_chk_host, 0);
      signal(SIGABRT, chk_abort_handler);
      atexit([] { if (g_chk_host && g_chk_host[0]) { signal(SIGABRT, SIG_IGN); chk_abort_handler(0); } });
    }
  });
  return dev;
}
#else
long long* chk_record() { return nullptr; }
#endif

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch() {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return MOT_OK;
  g_last_cuda_error = e;
  return MOT_ERR_CUDA;
}

int device_props(int* sm_count, int* smem_optin) {
  static std::mutex mu;
  static int cached_dev = -1, cached_sms = 0, cached_optin = 0, cached_major = 0;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    g_last_cuda_error = cudaGetLastError();
    return MOT_ERR_NO_DEVICE;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (dev != cached_dev) {
    int sms = 0, optin = 0, major = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      g_last_cuda_error = cudaGetLastError();
      return MOT_ERR_NO_DEVICE;
    }
    cached_dev = dev;
    cached_sms = sms;
    cached_optin = optin;
    cached_major = major;
  }
  if (cached_major != 10) return MOT_ERR_NO_DEVICE;  // sm_100a binary only
  *sm_count = cached_sms;
  *smem_optin = cached_optin;
  return MOT_OK;
}

}  // namespace mot

extern "C" const char* mot_strerror(int rc) {
  switch (rc) {
    case MOT_OK: return "ok";
    case MOT_ERR_BAD_ARG: return "bad argument (null pointer, size or inconsistent dims)";
    case MOT_ERR_UNSUPPORTED: return "unsupported option (this library has no fallback path)";
    case MOT_ERR_MISALIGNED: return "pointer not 16-byte aligned or dim not a multiple of 8";
    case MOT_ERR_WORKSPACE: return "workspace too small (see mot_embed_workspace_bytes)";
    case MOT_ERR_CUDA: return "CUDA launch failed (see mot_last_cuda_error)";
    case MOT_ERR_NO_DEVICE: return "no sm_100 (B200) device is current";
  }
  return "unknown error";
}

extern "C" int mot_abi_version(void) { return MOT_B200_ABI_VERSION; }

extern "C" const char* mot_last_cuda_error(void) {
  return mot::g_last_cuda_error == cudaSuccess ? "" : cudaGetErrorString(mot::g_last_cuda_error);
}

extern "C" int64_t mot_launch_count(void) { return mot::g_launches.load(std::memory_order_relaxed); }
extern "C" void mot_launch_count_reset(void) { mot::g_launches.store(0, std::memory_order_relaxed); }

extern "C" void mot_profile_events(void* fwd_start, void* fwd_stop, void* bwd_start, void* bwd_stop) {
  mot::g_prof_fwd_start = reinterpret_cast<cudaEvent_t>(fwd_start);
  mot::g_prof_fwd_stop = reinterpret_cast<cudaEvent_t>(fwd_stop);
  mot::g_prof_start = reinterpret_cast<cudaEvent_t>(bwd_start);
  mot::g_prof_stop = reinterpret_cast<cudaEvent_t>(bwd_stop);
}

extern "C" void mot_profile_trace(void* device_buffer) { mot::g_trace = reinterpret_cast<long long*>(device_buffer); }
