// Layout kernels: NCHW fp32 input -> NHWC im2col operand of the first conv; weight / weight-gradient
// permutations between PyTorch's OIHW / IOHW fp32 parameters and the K-major operand packs.
#include "common.cuh"

namespace unetk {

// out[n,h,w,k] = x[n, k % cin, h + (k/cin)/3 - 1, w + (k/cin)%3 - 1]  for k < 9*cin, else 0
template <typename T>
__global__ void __launch_bounds__(256) im2col_first_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                                           T* __restrict__ out, int kpad, int ld) {
  const int kg = kpad / 8;
  const int64_t total = (int64_t)n * h * w * kg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % kg);
    int64_t p = i / kg;
    const int px = (int)(p % w);
    const int py = (int)((p / w) % h);
    const int img = (int)(p / ((int64_t)w * h));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < 9 * cin) {
        const int t = k / cin, ci = k % cin;
        const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) val = __ldg(x + (((int64_t)img * cin + ci) * h + yy) * w + xx);
      }
      v[j] = val;
    }
    store8(out + p * ld + g * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) permute3_kernel(const float* __restrict__ src, T* __restrict__ dst, int d0, int d1,
                                                       int d2, int64_t ss0, int64_t ss1, int64_t ss2, int64_t ds0,
                                                       int64_t ds1, int64_t ds2) {
  // one thread per (i0, i1); the short dim i2 (taps) is looped
  const int64_t total = (int64_t)d0 * d1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = i / d1, i1 = i % d1;
    const float* s = src + i0 * ss0 + i1 * ss1;
    T* d = dst + i0 * ds0 + i1 * ds1;
    for (int i2 = 0; i2 < d2; ++i2) d[i2 * ds2] = from_f<T>(s[i2 * ss2]);
  }
}

}  // namespace unetk

using namespace unetk;

extern "C" {

int unetk_im2col3x3_first(const float* x_nchw, int32_t n, int32_t cin, int32_t h, int32_t w, const unetk_tensor* out,
                          void* stream) {
  UNETK_REQUIRE(x_nchw && out, "im2col_first: null argument");
  UNETK_REQUIRE(tensor_ok(*out) && vec8_ok(*out), "im2col_first: out must be NHWC with c%%8==0, ld%%8==0");
  UNETK_REQUIRE(out->n == n && out->h == h && out->w == w && out->c >= 9 * cin && cin > 0,
                "im2col_first: out must be [N,H,W,K>=9*Cin]");
  const int64_t items = (int64_t)n * h * w * (out->c / 8);
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  UNETK_DISPATCH_DTYPE(out->dtype, T, {
    im2col_first_kernel<T><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x_nchw, n, cin, h, w, (T*)out->ptr, out->c, out->ld);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_permute3(const float* src, void* dst, int32_t dst_dtype, int32_t d0, int32_t d1, int32_t d2, int64_t ss0,
                   int64_t ss1, int64_t ss2, int64_t ds0, int64_t ds1, int64_t ds2, void* stream) {
  UNETK_REQUIRE(src && dst && d0 > 0 && d1 > 0 && d2 > 0, "permute3: bad argument");
  UNETK_REQUIRE(dst_dtype == UNETK_F32 || dst_dtype == UNETK_BF16, "permute3: bad dtype");
  const int64_t items = (int64_t)d0 * d1;
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  UNETK_DISPATCH_DTYPE(dst_dtype, T, {
    permute3_kernel<T><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (T*)dst, d0, d1, d2, ss0, ss1, ss2, ds0, ds1, ds2);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}
}
