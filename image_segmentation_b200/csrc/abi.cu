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

int64_t unetk_query_workspace(int32_t what, int32_t a, int32_t b, int32_t c, int32_t d) {
  if (a <= 0) {
    unetk::set_error("query_workspace: first size must be positive");
    return UNETK_ERR_INVALID;
  }
  switch (what) {
    case UNETK_WS_BN_STATS:
    case UNETK_WS_BN_BWD_SUMS: return 2LL * a * (int64_t)sizeof(double);
    case UNETK_WS_HEAD_BN_SUMS: return b > 0 ? (3LL + b) * a * (int64_t)sizeof(double) : UNETK_ERR_INVALID;
    case UNETK_WS_POOL_IDX:
      if (b <= 0 || c <= 0 || d <= 0 || (b & 1) || (c & 1) || (d % 8)) break;
      return (int64_t)a * (b / 2) * (c / 2) * (d / 8) * 2;
    case UNETK_WS_WGRAD: {
      if (a > 2 || b <= 0 || c <= 0) break;
      const int taps = a == 0 ? 1 : (a == 1 ? 9 : 4);
      return (int64_t)b * taps * c * (int64_t)sizeof(float);
    }
    case UNETK_WS_DICE_ACCUM: return (3LL * a + 2) * (int64_t)sizeof(double);
    case UNETK_WS_DICE_COEF: return (2LL * a + 1) * (int64_t)sizeof(float);
    case UNETK_WS_EVAL_ACCUM: return b > 0 ? (int64_t)a * (3LL * b + 2) * (int64_t)sizeof(double) : UNETK_ERR_INVALID;
    case UNETK_WS_CONFUSION: return 4LL * a * (int64_t)sizeof(int64_t);
    default: break;
  }
  unetk::set_error("query_workspace: unknown buffer %d or bad sizes (%d,%d,%d,%d)", what, a, b, c, d);
  return UNETK_ERR_INVALID;
}

int32_t unetk_struct_size(int32_t which) {
  switch (which) {
    case 0: return (int32_t)sizeof(unetk_tensor);
    case 1: return (int32_t)sizeof(unetk_conv_args);
    case 2: return (int32_t)sizeof(unetk_wgrad_args);
    case 3: return (int32_t)sizeof(unetk_bn_finalize_args);
    case 4: return (int32_t)sizeof(unetk_bn_bwd_args);
    case 5: return (int32_t)sizeof(unetk_wjob);
    case 6: return (int32_t)sizeof(unetk_dice_ce_args);
    case 7: return (int32_t)sizeof(unetk_head_bn_bwd_args);
    case 8: return (int32_t)sizeof(unetk_eval_image);
    case 9: return (int32_t)sizeof(unetk_eval_args);
    default: return -1;
  }
}

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
