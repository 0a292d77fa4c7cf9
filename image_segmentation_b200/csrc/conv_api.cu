// unetk_conv / unetk_wgrad / unetk_channel_sum: argument validation and dispatch between the fp32 parity tier
// (CUDA-core kernels) and the bf16 throughput tier (TMA + tcgen05 kernels).  There is no CPU path.
#include <string.h>

#include "conv_internal.cuh"

namespace unetk {

template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* __restrict__ t, int64_t npix, int ld, int c,
                                                          int pix_per_block, float* __restrict__ out,
                                                          float* __restrict__ partial) {
  // block = 256/cgb pixel rows x cgb channel groups (8 channels each)
  extern __shared__ float red[];
  const int cg = c / 8;
  int cgb = 32;
  while (cgb > 1 && (cg % cgb) != 0) cgb >>= 1;
  const int rows = 256 / cgb;
  const int lane_g = threadIdx.x % cgb, row = threadIdx.x / cgb;
  const int g = blockIdx.x * cgb + lane_g;
  const int64_t p0 = (int64_t)blockIdx.y * pix_per_block, p1 = min(p0 + (int64_t)pix_per_block, npix);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t p = p0 + row; p < p1; p += rows) {
    float v[8];
    load8(t + p * ld + (size_t)g * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += v[i];
  }
  const int width = cgb * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) red[row * width + lane_g * 8 + i] = acc[i];
  __syncthreads();
  for (int ch = threadIdx.x; ch < width; ch += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += red[r * width + ch];
    if (partial)      // ordered mode: block partials are stored and summed in index order by channel_sum_finish_kernel
      partial[(size_t)blockIdx.y * c + (size_t)blockIdx.x * width + ch] = s;
    else
      atomicAdd(out + (size_t)blockIdx.x * width + ch, s);
  }
}

__global__ void __launch_bounds__(256) channel_sum_finish_kernel(const float* __restrict__ partial, int nblocks, int c,
                                                                 float* __restrict__ out) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * c + ch];
  out[ch] += s;
}

}  // namespace unetk

using namespace unetk;

extern "C" {

int unetk_conv(const unetk_conv_args* a, void* stream) {
  UNETK_REQUIRE(a != nullptr, "conv: null args");
  UNETK_REQUIRE(tensor_ok(a->x) && tensor_ok(a->y) && a->w, "conv: bad x/y/w");
  UNETK_REQUIRE(a->x.dtype == a->y.dtype, "conv: x and y must share a dtype");
  UNETK_REQUIRE(a->mode >= 0 && a->mode <= 3, "conv: mode must be 0..3");
  ConvGeom g;
  g.cin = a->x.c;
  g.cout = a->y.c;
  g.cout_total = a->y.c;
  g.taps = a->mode == 0 ? 1 : (a->mode == 1 ? 9 : (a->mode == 2 ? 1 : 4));
  if (a->mode <= 1) {
    UNETK_REQUIRE(a->y.n == a->x.n && a->y.h == a->x.h && a->y.w == a->x.w, "conv: y must have x's spatial shape");
    g.rows_n = a->y.n; g.rows_h = a->y.h; g.rows_w = a->y.w;
  } else if (a->mode == 2) {
    UNETK_REQUIRE(a->y.n == a->x.n && a->y.h == 2 * a->x.h && a->y.w == 2 * a->x.w, "conv(T fprop): y must be [N,2h,2w,Cout]");
    g.cout_total = 4 * a->y.c;
    g.rows_n = a->x.n; g.rows_h = a->x.h; g.rows_w = a->x.w;
  } else {
    UNETK_REQUIRE(a->y.n == a->x.n && 2 * a->y.h == a->x.h && 2 * a->y.w == a->x.w, "conv(T dgrad): x must be [N,2h,2w,C]");
    g.rows_n = a->y.n; g.rows_h = a->y.h; g.rows_w = a->y.w;
  }
  UNETK_REQUIRE((a->stat_sum == nullptr) == (a->stat_sumsq == nullptr), "conv: stat_sum and stat_sumsq come together");
  UNETK_REQUIRE(!(a->stat_sum && a->mode >= 2), "conv: statistics only for modes 0/1");
  const bool bnred = a->bn_z.ptr != nullptr;
  if (bnred) {
    UNETK_REQUIRE(a->mode != 2, "conv: fused BatchNorm-backward reduction is for data-gradient launches (modes 0/1/3)");
    UNETK_REQUIRE(tensor_ok(a->bn_z) && a->bn_z.dtype == a->y.dtype && a->bn_z.n == a->y.n && a->bn_z.h == a->y.h &&
                      a->bn_z.w == a->y.w && a->bn_z.c == a->y.c, "conv: bn_z must have the shape and dtype of y");
    UNETK_REQUIRE(a->bn_scale && a->bn_shift && a->bn_mean && a->bn_invstd && a->bn_sums, "conv: bn_* pointers missing");
  }

  int algo = a->algo & UNETK_ALGO_MASK;
  if (algo == UNETK_ALGO_AUTO) {
    // bf16 -> tensor cores whenever the shape allows it (channel counts multiples of 64); other bf16 shapes (e.g. a
    // 32-channel layer) run on the CUDA-core kernels -- both are this library's own kernels, neither is a fallback to
    // another backend.  UNETK_ALGO_TC requests the tcgen05 path explicitly and fails loudly when it cannot run.
    const char* why = "";
    algo = (a->x.dtype == UNETK_BF16 && tc_conv_supported(a, g, &why)) ? UNETK_ALGO_TC : UNETK_ALGO_SIMT;
  }
  if (algo == UNETK_ALGO_TC) {
    const char* why = "";
    if (!tc_conv_supported(a, g, &why)) {
      set_error("conv: tcgen05 path does not support this problem: %s", why);
      return UNETK_ERR_UNSUPPORTED;
    }
    return tc_conv(a, g, (cudaStream_t)stream);
  }
  UNETK_REQUIRE(algo == UNETK_ALGO_SIMT, "conv: unknown algo %d", algo);
  int rc = simt_conv(a, g, (cudaStream_t)stream);
  if (rc) return rc;
  if (a->stat_sum && (rc = unetk_bn_stats(&a->y, a->stat_sum, a->stat_sumsq, stream))) return rc;
  if (bnred) {
    // CUDA-core tier: same contract, realised with the stand-alone reduction kernel
    unetk_bn_bwd_args b;
    memset(&b, 0, sizeof(b));
    b.z = a->bn_z;
    b.dy = a->y;
    b.scale = a->bn_scale; b.shift = a->bn_shift; b.mean = a->bn_mean; b.invstd = a->bn_invstd;
    b.sums = a->bn_sums;
    return unetk_bn_relu_bwd_reduce(&b, stream);
  }
  return UNETK_OK;
}

int unetk_wgrad(const unetk_wgrad_args* a, void* stream) {
  UNETK_REQUIRE(a != nullptr, "wgrad: null args");
  UNETK_REQUIRE(tensor_ok(a->u) && tensor_ok(a->s) && a->dw, "wgrad: bad u/s/dw");
  UNETK_REQUIRE(a->u.dtype == a->s.dtype, "wgrad: u and s must share a dtype");
  UNETK_REQUIRE(a->mode >= 0 && a->mode <= 2, "wgrad: mode must be 0..2");
  const int taps = a->mode == 0 ? 1 : (a->mode == 1 ? 9 : 4);
  if (a->mode <= 1)
    UNETK_REQUIRE(a->s.n == a->u.n && a->s.h == a->u.h && a->s.w == a->u.w, "wgrad: u and s must share a spatial shape");
  else
    UNETK_REQUIRE(a->s.n == a->u.n && a->s.h == 2 * a->u.h && a->s.w == 2 * a->u.w, "wgrad(mode 2): s must be [N,2h,2w,C]");
  int algo = a->algo & UNETK_ALGO_MASK;
  if (algo == UNETK_ALGO_AUTO) {
    const char* why0 = "";
    algo = (a->u.dtype == UNETK_BF16 && tc_wgrad_supported(a, taps, &why0)) ? UNETK_ALGO_TC : UNETK_ALGO_SIMT;
  }
  if (algo == UNETK_ALGO_TC) {
    const char* why = "";
    if (!tc_wgrad_supported(a, taps, &why)) {
      set_error("wgrad: tcgen05 path does not support this problem: %s", why);
      return UNETK_ERR_UNSUPPORTED;
    }
    return tc_wgrad(a, taps, (cudaStream_t)stream);
  }
  UNETK_REQUIRE(algo == UNETK_ALGO_SIMT, "wgrad: unknown algo %d", algo);
  return simt_wgrad(a, taps, (cudaStream_t)stream);
}

int64_t unetk_wgrad_partial_bytes(const unetk_wgrad_args* a) {
  if (!a || !tensor_ok(a->u) || !tensor_ok(a->s) || a->mode < 0 || a->mode > 2) {
    set_error("wgrad_partial_bytes: bad arguments");
    return UNETK_ERR_INVALID;
  }
  int algo = a->algo & UNETK_ALGO_MASK;
  if (algo == UNETK_ALGO_AUTO) algo = a->u.dtype == UNETK_BF16 ? UNETK_ALGO_TC : UNETK_ALGO_SIMT;
  const char* why = "";
  if (algo != UNETK_ALGO_TC || !tc_wgrad_supported(a, 0, &why)) return 0;   // the CUDA-core tier needs no partial buffer
  return tc_wgrad_partial_bytes(a, a->mode == 0 ? 1 : (a->mode == 1 ? 9 : 4));
}

static int channel_sum_impl(const unetk_tensor* t, float* out, float* scratch, int64_t scratch_bytes, void* stream);

int unetk_channel_sum(const unetk_tensor* t, float* out, void* stream) { return channel_sum_impl(t, out, nullptr, 0, stream); }

int unetk_channel_sum_ordered(const unetk_tensor* t, float* out, float* scratch, int64_t scratch_bytes, void* stream) {
  UNETK_REQUIRE(scratch != nullptr, "channel_sum_ordered: scratch buffer missing");
  return channel_sum_impl(t, out, scratch, scratch_bytes, stream);
}

static int channel_sum_impl(const unetk_tensor* t, float* out, float* scratch, int64_t scratch_bytes, void* stream) {
  UNETK_REQUIRE(t && out, "channel_sum: null argument");
  UNETK_REQUIRE(tensor_ok(*t) && vec8_ok(*t), "channel_sum: t must be NHWC with c%%8==0, ld%%8==0, 16B aligned");
  const int cg = t->c / 8;
  int cgb = 32;
  while (cgb > 1 && (cg % cgb) != 0) cgb >>= 1;
  const int rows = 256 / cgb;
  const int64_t npix = pixels(*t);
  int ppb = rows * 32;
  if ((npix + ppb - 1) / ppb > 65535) ppb *= 16;
  if (scratch) {
    // ordered mode: at most UNETK_CHANNEL_SUM_BLOCKS block rows, whose partial sums fit the caller's scratch
    while ((npix + ppb - 1) / ppb > UNETK_CHANNEL_SUM_BLOCKS) ppb *= 2;
    UNETK_REQUIRE(scratch_bytes >= (int64_t)UNETK_CHANNEL_SUM_BLOCKS * t->c * (int64_t)sizeof(float),
                  "channel_sum_ordered: scratch must hold UNETK_CHANNEL_SUM_BLOCKS * C floats");
  }
  UNETK_REQUIRE((npix + ppb - 1) / ppb <= 65535, "channel_sum: tensor too large");
  dim3 grid(cg / cgb, (unsigned)((npix + ppb - 1) / ppb));
  const size_t smem = (size_t)rows * cgb * 8 * sizeof(float);
  UNETK_DISPATCH_DTYPE(t->dtype, T, {
    channel_sum_kernel<T><<<grid, 256, smem, (cudaStream_t)stream>>>((const T*)t->ptr, npix, t->ld, t->c, ppb, out, scratch);
  });
  UNETK_LAUNCH_CHECK();
  if (scratch) {
    channel_sum_finish_kernel<<<(t->c + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scratch, (int)grid.y, t->c, out);
    UNETK_LAUNCH_CHECK();
  }
  return UNETK_OK;
}
}
