// Library-level entry points of libgngf_sm100.so: status strings, device info, launch accounting.
#include <cstdlib>
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace gngf {
static std::atomic<int64_t> g_launches{0};
static thread_local char g_last_cuda_error[256] = "";

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
    // measurement aid (profiles/sm_limit_probe.py): persistent kernels launch on fewer SMs -- does a kernel's time scale
    // with the SMs it runs on (bound inside the SM) or stay put (bound by the L2 / fabric they share)?
    if (const char* lim = getenv("GNGF_DEBUG_SM_LIMIT")) {
      const int v = atoi(lim);
      if (v > 0 && v < cached) cached = v;
    }
  }
  return cached;
}

int debug_no_skip() {
  const char* v = getenv("GNGF_DEBUG_NO_SKIP");
  return (v && v[0] == '1') ? 1 : 0;
}

int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return GNGF_OK;
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "CUDA error: %s", cudaGetErrorString(e));
  return GNGF_ERR_CUDA;
}
}  // namespace gngf

extern "C" {

const char* gngf_strerror(int status) {
  switch (status) {
    case GNGF_OK: return "ok";
    case GNGF_ERR_INVALID_ARGUMENT: return "invalid argument";
    case GNGF_ERR_UNSUPPORTED: return "unsupported configuration";
    case GNGF_ERR_CUDA: return gngf::g_last_cuda_error[0] ? gngf::g_last_cuda_error : "CUDA error";
    case GNGF_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
  }
}

int gngf_abi_version(void) { return 1; }

int gngf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return GNGF_ERR_NO_DEVICE;
  }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    cudaGetLastError();
    return GNGF_ERR_NO_DEVICE;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return GNGF_OK;
}

int64_t gngf_launch_count(void) { return gngf::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
