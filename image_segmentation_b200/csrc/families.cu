// Bandwidth kernels of the other model families that share the U-Net's blocks:
//
//   unetk_bias_sigmoid_fwd / _bwd   autoencoder/autoencoder.py:188-191   Conv3x3(64 -> dout) bias + nn.Sigmoid of the reconstruction
//                                   output (the conv itself runs on the tensor cores with Cout zero-padded to 64)
//   unetk_bilinear_up_fwd / _bwd    clip/clipunet.py:99-100              F.interpolate(skip, size, 'bilinear', align_corners=False)
//                                   in NHWC, written straight into the concat slice; backward as a gather (no atomics)
//   unetk_prompt_compose_fwd / _bwd prompt_based/prompt.py:33-56         softmax(clip) x sigmoid(mask) probability composition
//   unetk_nchw_to_nhwc / unetk_nhwc_to_nchw                              layout converters of the block-level API
//                                   (DoubleConvReLU / Down / Up called on their own, unet/unet.py:24,44,62)
//
// Roofline: HBM; each tensor is read once and written once at its storage type.
#include "common.cuh"

namespace unetk {

// ------------------------------------------------------------------------------------------------------------------
// bias + sigmoid, NHWC (first `dout` of z.c channels) -> NCHW fp32
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bias_sigmoid_fwd_kernel(const T* __restrict__ z, int ld, int64_t npix, int64_t hw,
                                                               const float* __restrict__ bias, int dout,
                                                               float* __restrict__ out) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    load8(z + p * ld, v);
    const int64_t img = p / hw, off = p % hw;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < dout) {
        const float x = v[k] + (bias ? __ldg(bias + k) : 0.f);
        out[(img * dout + k) * hw + off] = 1.f / (1.f + expf(-x));
      }
    }
  }
}

// dz[p, k] = dy[p, k] * s (1 - s) for k < dout, 0 for the 8 - dout padding channels of the first 8-channel group (the
// remaining channels of dz are zeroed once by the caller and never written); db[k] += sum_p dz[p, k]
template <typename T>
__global__ void __launch_bounds__(256) bias_sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ s,
                                                               int64_t npix, int64_t hw, int dout, T* __restrict__ dz, int ld,
                                                               float* __restrict__ db) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = p / hw, off = p % hw;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = 0.f;
      if (k < dout) {
        const int64_t i = (img * dout + k) * hw + off;
        const float sv = s[i];
        v[k] = round_to<T>(dy[i] * sv * (1.f - sv));
        acc[k] += v[k];
      }
    }
    store8(dz + p * ld, v);
  }
  if (db) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < dout) {
        const float r = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(db + k, r);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// bilinear up-sampling, NHWC, align_corners = False (ATen's area_pixel_compute_source_index: max(0, scale*(dst+0.5)-0.5))
// ------------------------------------------------------------------------------------------------------------------
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_of(int dst, float scale, int in_size) {
  float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
  src = src < 0.f ? 0.f : src;
  Lerp r;
  r.i0 = (int)src;
  if (r.i0 > in_size - 1) r.i0 = in_size - 1;
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

// forward: one block per destination row (img, y) -- no 64-bit index arithmetic in the inner loop, the row's vertical taps
// are computed once; threads walk the row's (x, 8-channel group) items with consecutive threads on consecutive 16 bytes
template <typename T>
__global__ void __launch_bounds__(256) bilinear_up_fwd_kernel(const T* __restrict__ src, int sld, int n, int ih, int iw, int c,
                                                              T* __restrict__ dst, int dld, int oh, int ow, float sh, float sw) {
  const uint32_t cg = (uint32_t)c / 8;
  const uint32_t row = blockIdx.x;                       // img * oh + y
  const uint32_t y = row % (uint32_t)oh, img = row / (uint32_t)oh;
  const Lerp ly = lerp_of((int)y, sh, ih);
  const T* base0 = src + ((int64_t)img * ih + ly.i0) * iw * (int64_t)sld;
  const T* base1 = src + ((int64_t)img * ih + ly.i1) * iw * (int64_t)sld;
  T* drow = dst + (int64_t)row * ow * (int64_t)dld;
  const uint32_t items = (uint32_t)ow * cg;
  for (uint32_t i = threadIdx.x; i < items; i += blockDim.x) {
    const uint32_t x = i / cg, g = i - x * cg;
    const Lerp lx = lerp_of((int)x, sw, iw);
    float a[8], b[8], cc[8], d[8], o[8];
    load8(base0 + (int64_t)lx.i0 * sld + g * 8, a);
    load8(base0 + (int64_t)lx.i1 * sld + g * 8, b);
    load8(base1 + (int64_t)lx.i0 * sld + g * 8, cc);
    load8(base1 + (int64_t)lx.i1 * sld + g * 8, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = ly.l0 * (lx.l0 * a[k] + lx.l1 * b[k]) + ly.l1 * (lx.l0 * cc[k] + lx.l1 * d[k]);
    store8(drow + (int64_t)x * dld + g * 8, o);
  }
}

// gather form of the transpose (no atomics): a warp owns one source pixel and `cgl` = min(C/8, 32) consecutive 8-channel
// groups; its 32 / cgl lane rows share the walk over the destination window whose interpolation footprint contains that
// source pixel ((2/scale + 2)^2 positions: 34 x 34 for the 16x up-sampling of the last decoder block, where C/8 = 8 and
// four lane rows split every window row) and are summed with shuffles.  Lanes of one lane row read consecutive 16-byte
// pieces of one pixel; the vertical weight is evaluated once per window row.
template <typename T>
__global__ void __launch_bounds__(256) bilinear_up_bwd_kernel(const T* __restrict__ dd, int dld, int n, int oh, int ow, int c,
                                                              T* __restrict__ ds, int sld, int ih, int iw, float sh, float sw,
                                                              int cgl) {
  const int cg = c / 8;
  const int chunks = cg / cgl;                       // cgl: largest power of two <= 32 dividing cg
  const int rows = 32 / cgl;                         // lane rows sharing the window walk
  const int lane = threadIdx.x & 31;
  const int gl = lane % cgl, pr = lane / cgl;
  const int64_t warps_total = (int64_t)n * ih * iw * chunks;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wi = warp0; wi < warps_total; wi += wstride) {
    const int chunk = (int)(wi % chunks);
    const int64_t p = wi / chunks;
    const int sx = (int)(p % iw), sy = (int)((p / iw) % ih);
    const int64_t img = p / ((int64_t)iw * ih);
    const int g = chunk * cgl + gl;
    // rows y whose source coordinate lies in (sy - 1, sy + 1): y in ((sy - 0.5)/sh - 0.5, (sy + 1.5)/sh - 0.5); one row of
    // slack on either side (border clamping is handled by lerp_of itself: weights are evaluated, never assumed)
    int y0 = (int)floorf(((float)sy - 0.5f) / sh - 0.5f) - 1, y1 = (int)ceilf(((float)sy + 1.5f) / sh - 0.5f) + 1;
    int x0 = (int)floorf(((float)sx - 0.5f) / sw - 0.5f) - 1, x1 = (int)ceilf(((float)sx + 1.5f) / sw - 0.5f) + 1;
    y0 = y0 < 0 ? 0 : y0; x0 = x0 < 0 ? 0 : x0;
    y1 = y1 > oh - 1 ? oh - 1 : y1; x1 = x1 > ow - 1 ? ow - 1 : x1;
    if (sy == 0) y0 = 0;                              // negative source coordinates clamp to row / column 0
    if (sx == 0) x0 = 0;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const T* base = dd + img * oh * ow * (int64_t)dld + g * 8;
    for (int y = y0; y <= y1; ++y) {
      const Lerp ly = lerp_of(y, sh, ih);
      const float wy = (ly.i0 == sy ? ly.l0 : 0.f) + (ly.i1 == sy ? ly.l1 : 0.f);
      if (wy == 0.f) continue;                        // warp-uniform
      const T* rowp = base + (int64_t)y * ow * dld;
      for (int x = x0 + pr; x <= x1; x += rows) {
        const Lerp lx = lerp_of(x, sw, iw);
        const float wgt = wy * ((lx.i0 == sx ? lx.l0 : 0.f) + (lx.i1 == sx ? lx.l1 : 0.f));
        if (wgt != 0.f) {
          float v[8];
          load8(rowp + (int64_t)x * dld, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, v[k], acc[k]);
        }
      }
    }
    for (int o = cgl; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (pr == 0) store8(ds + p * sld + g * 8, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// prompt composition (prompt_based/prompt.py:33-56), NCHW fp32:
//   p = softmax(clip_logit[4]), m = sigmoid(mask_logit);  final = [1 - m, m (p0 + p3), m p1, m p2]
//   d mask_logit = m (1 - m) * (-g0 + g1 (p0 + p3) + g2 p1 + g3 p2)         (the CLIP branch is frozen: no gradient)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void softmax4(const float* __restrict__ x, int64_t base, int64_t hw, float (&p)[4]) {
  float v[4], m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = x[base + k * hw];
    m = fmaxf(m, v[k]);
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    p[k] = expf(v[k] - m);
    sum += p[k];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] *= inv;
}

__global__ void __launch_bounds__(256) prompt_compose_fwd_kernel(const float* __restrict__ clip, const float* __restrict__ mask,
                                                                 int64_t npix, int64_t hw, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = i / hw, off = i % hw, base = img * 4 * hw + off;
    float p[4];
    softmax4(clip, base, hw, p);
    const float m = 1.f / (1.f + expf(-mask[i]));
    // same operation order as the reference: selected = m * p; final[1] = selected[0] + selected[3]
    out[base] = 1.f - m;
    out[base + hw] = m * p[0] + m * p[3];
    out[base + 2 * hw] = m * p[1];
    out[base + 3 * hw] = m * p[2];
  }
}

__global__ void __launch_bounds__(256) prompt_compose_bwd_kernel(const float* __restrict__ clip, const float* __restrict__ mask,
                                                                 const float* __restrict__ g, int64_t npix, int64_t hw,
                                                                 float* __restrict__ dmask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = i / hw, off = i % hw, base = img * 4 * hw + off;
    float p[4];
    softmax4(clip, base, hw, p);
    const float m = 1.f / (1.f + expf(-mask[i]));
    const float dm = -g[base] + g[base + hw] * (p[0] + p[3]) + g[base + 2 * hw] * p[1] + g[base + 3 * hw] * p[2];
    dmask[i] = dm * m * (1.f - m);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// NCHW fp32 <-> NHWC T, 32 x 32 (channel x pixel) shared-memory tiles
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, int c, int64_t hw, T* __restrict__ dst,
                                                           int ld) {
  __shared__ float tile[32][33];
  const int64_t img = blockIdx.z, p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int ch = c0 + r;
    const int64_t p = p0 + tx;
    tile[r][tx] = (ch < c && p < hw) ? src[(img * c + ch) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int ch = c0 + tx;
    if (p < hw && ch < c) dst[(img * hw + p) * ld + ch] = from_f<T>(tile[tx][r]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ src, int ld, int c, int64_t hw,
                                                           float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int64_t img = blockIdx.z, p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int ch = c0 + tx;
    tile[r][tx] = (p < hw && ch < c) ? to_f(src[(img * hw + p) * ld + ch]) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ch = c0 + r;
    const int64_t p = p0 + tx;
    if (ch < c && p < hw) dst[(img * c + ch) * hw + p] = tile[tx][r];
  }
}

static int grid_1d(int64_t items, int threads, int per_sm) {
  int64_t blocks = (items + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  return (int)(blocks > 0 ? blocks : 1);
}

}  // namespace unetk

using namespace unetk;

extern "C" {

int unetk_bias_sigmoid_fwd(const unetk_tensor* z, const float* bias, int32_t dout, float* out_nchw, void* stream) {
  UNETK_REQUIRE(z && out_nchw, "bias_sigmoid_fwd: null argument");
  UNETK_REQUIRE(tensor_ok(*z) && vec8_ok(*z), "bias_sigmoid_fwd: z must be NHWC with c%%8==0, ld%%8==0, 16B aligned");
  UNETK_REQUIRE(dout >= 1 && dout <= 8, "bias_sigmoid_fwd: 1..8 output channels supported");
  const int64_t npix = pixels(*z), hw = (int64_t)z->h * z->w;
  UNETK_DISPATCH_DTYPE(z->dtype, T, {
    bias_sigmoid_fwd_kernel<T><<<grid_1d(npix, 256, 8), 256, 0, (cudaStream_t)stream>>>((const T*)z->ptr, z->ld, npix, hw, bias, dout, out_nchw);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_bias_sigmoid_bwd(const float* dy_nchw, const float* out_nchw, int32_t dout, const unetk_tensor* dz, float* dbias,
                           void* stream) {
  UNETK_REQUIRE(dy_nchw && out_nchw && dz, "bias_sigmoid_bwd: null argument");
  UNETK_REQUIRE(tensor_ok(*dz) && vec8_ok(*dz), "bias_sigmoid_bwd: dz must be NHWC with c%%8==0, ld%%8==0, 16B aligned");
  UNETK_REQUIRE(dout >= 1 && dout <= 8, "bias_sigmoid_bwd: 1..8 output channels supported");
  const int64_t npix = pixels(*dz), hw = (int64_t)dz->h * dz->w;
  UNETK_DISPATCH_DTYPE(dz->dtype, T, {
    bias_sigmoid_bwd_kernel<T><<<grid_1d(npix, 256, 8), 256, 0, (cudaStream_t)stream>>>(dy_nchw, out_nchw, npix, hw, dout, (T*)dz->ptr, dz->ld, dbias);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

static int check_bilinear(const unetk_tensor* lo, const unetk_tensor* hi, const char* who) {
  UNETK_REQUIRE(lo && hi, "%s: null argument", who);
  UNETK_REQUIRE(tensor_ok(*lo) && vec8_ok(*lo) && tensor_ok(*hi) && vec8_ok(*hi), "%s: tensors must be NHWC with c%%8==0, ld%%8==0", who);
  UNETK_REQUIRE(lo->dtype == hi->dtype && lo->n == hi->n && lo->c == hi->c, "%s: batch / channels / dtype must match", who);
  return UNETK_OK;
}

int unetk_bilinear_up_fwd(const unetk_tensor* src, const unetk_tensor* dst, void* stream) {
  int rc = check_bilinear(src, dst, "bilinear_up_fwd");
  if (rc) return rc;
  const float sh = (float)src->h / (float)dst->h, sw = (float)src->w / (float)dst->w;
  const int64_t rows = (int64_t)dst->n * dst->h;
  UNETK_REQUIRE(rows < (1LL << 31) && (int64_t)dst->w * (dst->c / 8) < (1LL << 31), "bilinear_up_fwd: tensor too large");
  UNETK_DISPATCH_DTYPE(src->dtype, T, {
    bilinear_up_fwd_kernel<T><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(
        (const T*)src->ptr, src->ld, src->n, src->h, src->w, src->c, (T*)dst->ptr, dst->ld, dst->h, dst->w, sh, sw);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_bilinear_up_bwd(const unetk_tensor* ddst, const unetk_tensor* dsrc, void* stream) {
  int rc = check_bilinear(dsrc, ddst, "bilinear_up_bwd");
  if (rc) return rc;
  const float sh = (float)dsrc->h / (float)ddst->h, sw = (float)dsrc->w / (float)ddst->w;
  const int cg = dsrc->c / 8;
  int cgl = 1;
  while (cgl < 32 && cg % (cgl * 2) == 0) cgl *= 2;   // largest power of two <= 32 dividing the number of channel groups
  const int64_t warps = pixels(*dsrc) * (cg / cgl);
  UNETK_DISPATCH_DTYPE(dsrc->dtype, T, {
    bilinear_up_bwd_kernel<T><<<grid_1d(warps * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const T*)ddst->ptr, ddst->ld, ddst->n, ddst->h, ddst->w, ddst->c, (T*)dsrc->ptr, dsrc->ld, dsrc->h, dsrc->w, sh, sw, cgl);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_prompt_compose_fwd(const float* clip_logits, const float* mask_logits, int32_t n, int32_t h, int32_t w,
                             float* final_probs, void* stream) {
  UNETK_REQUIRE(clip_logits && mask_logits && final_probs && n > 0 && h > 0 && w > 0, "prompt_compose_fwd: bad argument");
  const int64_t hw = (int64_t)h * w, npix = (int64_t)n * hw;
  prompt_compose_fwd_kernel<<<grid_1d(npix, 256, 8), 256, 0, (cudaStream_t)stream>>>(clip_logits, mask_logits, npix, hw, final_probs);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_prompt_compose_bwd(const float* clip_logits, const float* mask_logits, const float* dfinal, int32_t n, int32_t h,
                             int32_t w, float* dmask_logits, void* stream) {
  UNETK_REQUIRE(clip_logits && mask_logits && dfinal && dmask_logits && n > 0 && h > 0 && w > 0, "prompt_compose_bwd: bad argument");
  const int64_t hw = (int64_t)h * w, npix = (int64_t)n * hw;
  prompt_compose_bwd_kernel<<<grid_1d(npix, 256, 8), 256, 0, (cudaStream_t)stream>>>(clip_logits, mask_logits, dfinal, npix, hw, dmask_logits);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_nchw_to_nhwc(const float* src_nchw, const unetk_tensor* dst, void* stream) {
  UNETK_REQUIRE(src_nchw && dst && tensor_ok(*dst), "nchw_to_nhwc: bad argument");
  const int64_t hw = (int64_t)dst->h * dst->w;
  UNETK_REQUIRE((hw + 31) / 32 < (1LL << 31) && dst->n <= 65535 && (dst->c + 31) / 32 <= 65535, "nchw_to_nhwc: tensor too large");
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((dst->c + 31) / 32), (unsigned)dst->n);
  UNETK_DISPATCH_DTYPE(dst->dtype, T, {
    nchw_to_nhwc_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(src_nchw, dst->c, hw, (T*)dst->ptr, dst->ld);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_nhwc_to_nchw(const unetk_tensor* src, float* dst_nchw, void* stream) {
  UNETK_REQUIRE(dst_nchw && src && tensor_ok(*src), "nhwc_to_nchw: bad argument");
  const int64_t hw = (int64_t)src->h * src->w;
  UNETK_REQUIRE((hw + 31) / 32 < (1LL << 31) && src->n <= 65535 && (src->c + 31) / 32 <= 65535, "nhwc_to_nchw: tensor too large");
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((src->c + 31) / 32), (unsigned)src->n);
  UNETK_DISPATCH_DTYPE(src->dtype, T, {
    nhwc_to_nchw_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)src->ptr, src->ld, src->c, hw, dst_nchw);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}
}
