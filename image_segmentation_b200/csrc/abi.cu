// Library-level entry points: version, thread-local error string, device query.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace unetk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  // immutable per-device attribute cache (the only global state of the library)
  static int cached[64];
  static std::once_flag once;
  std::call_once(once, [] {
    for (int i = 0; i < 64; ++i) cached[i] = 0;
  });
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace unetk

extern "C" {

int unetk_version(void) { return 100; }

const char* unetk_last_error(void) { return unetk::g_err; }

int unetk_device_query(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  UNETK_CUDA(cudaGetDevice(&dev));
  int n = 0, maj = 0, min = 0;
  UNETK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  UNETK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  UNETK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return UNETK_OK;
}
}
