// Shared helpers for libunetk.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/unetk.h"

namespace unetk {

void set_error(const char* fmt, ...);
int sm_count();

#define UNETK_REQUIRE(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      unetk::set_error(__VA_ARGS__);          \
      return UNETK_ERR_INVALID;               \
    }                                         \
  } while (0)

#define UNETK_CUDA(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      unetk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return UNETK_ERR_CUDA;                                                               \
    }                                                                                      \
  } while (0)

#define UNETK_LAUNCH_CHECK() UNETK_CUDA(cudaGetLastError())

inline bool tensor_ok(const unetk_tensor& t) {
  return t.ptr != nullptr && t.n > 0 && t.h > 0 && t.w > 0 && t.c > 0 && t.ld >= t.c &&
         (t.dtype == UNETK_F32 || t.dtype == UNETK_BF16);
}
inline bool vec8_ok(const unetk_tensor& t) {
  // 8-channel vector access: channel count, pixel stride and base pointer must keep 16 B alignment
  const size_t es = t.dtype == UNETK_BF16 ? 2 : 4;
  return (t.c % 8 == 0) && (t.ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(t.ptr) % 16) == 0) && es;
}
inline int64_t pixels(const unetk_tensor& t) { return (int64_t)t.n * t.h * t.w; }

// ---- element access -----------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
// value as it will be read back after being stored as T
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f(from_f<T>(v)); }

__device__ __forceinline__ void load8(const float* p, float v[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float v[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(u[i] << 16);
    v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float v[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]);
  r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]);
  r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void load4(const float* p, float v[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float v[4]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(r.x << 16);
  v[1] = __uint_as_float(r.x & 0xffff0000u);
  v[2] = __uint_as_float(r.y << 16);
  v[3] = __uint_as_float(r.y & 0xffff0000u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dispatch a templated launcher on dtype
#define UNETK_DISPATCH_DTYPE(dtype, T, ...)            \
  do {                                                 \
    if ((dtype) == UNETK_BF16) {                       \
      using T = __nv_bfloat16;                         \
      __VA_ARGS__                                      \
    } else {                                           \
      using T = float;                                 \
      __VA_ARGS__                                      \
    }                                                  \
  } while (0)

}  // namespace unetk
