// Evaluation tail: crop the padding off the network output, resize it back to each image's original size and --
// without materialising the resized logits -- reduce the per-image Dice+CE loss sums and the argmax confusion counts.
//
// Reference call sites replaced:
//   utils/utils.py:51-75, 101-115     reverse_resize_and_padding / process_batch_reverse (slice + F.interpolate per image)
//   utils/training.py:93-101          per image: loss_fn(pred[None], label[None]).item(), agg.accumulate(pred, label)
//
// The reference does, per image: 1 interpolate, ~14 loss kernels, ~10 metric kernels and 5 host syncs.  Here a batch
// of images of different sizes is one launch (blockIdx.y = image) plus a tiny finalize launch, and nothing is read
// back until the epoch ends.
//
// Interpolation arithmetic (ATen upsample_bilinear2d / upsample_nearest2d, align_corners=False), one rounding per
// operation so the CPU oracle (oracle/eval_oracle.py) can be matched bit for bit:
//   scale = float(in) / float(out);  src = max(fma(scale, dst + 0.5, -0.5), 0);  i0 = int(src);  i1 = i0 + (i0 < in-1)
//   l1 = src - i0;  l0 = 1 - l1;     out = ly0 * (lx0*v00 + lx1*v01) + ly1 * (lx0*v10 + lx1*v11)
//
// Roofline: HBM/L2 gather.  Algorithmic bytes per OUTPUT pixel: label (1 or 8) + 4*C when the resized logits are
// written; the source logits (4*C*crop_h*crop_w per image) are read once from HBM and re-read from L2.
#include "loss_common.cuh"

static_assert(sizeof(unetk_eval_image) == 32 && sizeof(unetk_eval_args) == 136, "ABI layout (see _lib.py)");

namespace unetk {

struct Tap {
  int i0, i1;
  float l0, l1;
};

__device__ __forceinline__ Tap bilinear_tap(int dst, int n_in, float scale) {
  float src = __fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f);
  src = fmaxf(src, 0.f);
  Tap t;
  t.i0 = min((int)src, n_in - 1);
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
  t.l1 = __fsub_rn(src, (float)t.i0);
  t.l0 = __fsub_rn(1.f, t.l1);
  return t;
}

__device__ __forceinline__ int nearest_tap(int dst, int n_in, float scale) {
  return min((int)floorf(__fmul_rn((float)dst, scale)), n_in - 1);
}

// x[k] = resized logit of class k at output pixel (oy, ox) of one image; plane = first class plane of the image's
// crop origin, hw = TH*TW
template <int MODE>
__device__ __forceinline__ void resized_pixel(const float* __restrict__ plane, int64_t hw, int tw, int c,
                                              const unetk_eval_image& im, float sy, float sx, int oy, int ox,
                                              float (&x)[kMaxClasses]) {
  if (MODE == 1) {
    const int64_t o = (int64_t)nearest_tap(oy, im.crop_h, sy) * tw + nearest_tap(ox, im.crop_w, sx);
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) x[k] = k < c ? plane[k * hw + o] : -INFINITY;
    return;
  }
  const Tap ty = bilinear_tap(oy, im.crop_h, sy), tx = bilinear_tap(ox, im.crop_w, sx);
  const int64_t r0 = (int64_t)ty.i0 * tw, r1 = (int64_t)ty.i1 * tw;
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) {
    if (k < c) {
      const float* q = plane + k * hw;
      const float v00 = q[r0 + tx.i0], v01 = q[r0 + tx.i1], v10 = q[r1 + tx.i0], v11 = q[r1 + tx.i1];
      const float top = __fadd_rn(__fmul_rn(tx.l0, v00), __fmul_rn(tx.l1, v01));
      const float bot = __fadd_rn(__fmul_rn(tx.l0, v10), __fmul_rn(tx.l1, v11));
      x[k] = __fadd_rn(__fmul_rn(ty.l0, top), __fmul_rn(ty.l1, bot));
    } else {
      x[k] = -INFINITY;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) crop_resize_kernel(const float* __restrict__ src, int c, int th, int tw,
                                                          const unetk_eval_image* __restrict__ images,
                                                          float* __restrict__ out) {
  const unetk_eval_image im = images[blockIdx.y];
  const int64_t hw = (int64_t)th * tw;
  const int npix = im.out_h * im.out_w;
  const float sy = __fdiv_rn((float)im.crop_h, (float)im.out_h), sx = __fdiv_rn((float)im.crop_w, (float)im.out_w);
  const float* plane = src + (int64_t)blockIdx.y * c * hw + (int64_t)im.crop_top * tw + im.crop_left;
  float* o = out + im.offset * c;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    float x[kMaxClasses];
    resized_pixel<MODE>(plane, hw, tw, c, im, sy, sx, p / im.out_w, p % im.out_w, x);
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) o[(int64_t)k * npix + p] = x[k];
  }
}

template <typename LabelT>
__global__ void __launch_bounds__(256) eval_reduce_kernel(unetk_eval_args a) {
  __shared__ unsigned int cm[kMaxClasses * kMaxClasses];
  __shared__ float red[8][3 * kMaxClasses + 2];
  for (int i = threadIdx.x; i < kMaxClasses * kMaxClasses; i += blockDim.x) cm[i] = 0;
  __syncthreads();
  const int c = a.c;
  const unetk_eval_image im = a.images[blockIdx.y];
  const int64_t hw = (int64_t)a.th * a.tw;
  const int npix = im.out_h * im.out_w;
  const float sy = __fdiv_rn((float)im.crop_h, (float)im.out_h), sx = __fdiv_rn((float)im.crop_w, (float)im.out_w);
  const float* plane = a.logits + (int64_t)blockIdx.y * c * hw + (int64_t)im.crop_top * a.tw + im.crop_left;
  const LabelT* label = static_cast<const LabelT*>(a.labels) + im.offset;
  float I[kMaxClasses], P[kMaxClasses], G[kMaxClasses], cen = 0.f, ced = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) I[k] = P[k] = G[k] = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const int64_t yl = (int64_t)label[p];
    if (yl < 0 || yl >= c) {
      atomicOr(a.status, 1);
      continue;
    }
    const int y = (int)yl;
    float x[kMaxClasses];
    resized_pixel<0>(plane, hw, a.tw, c, im, sy, sx, p / im.out_w, p % im.out_w, x);
    atomicAdd(&cm[y * kMaxClasses + argmax_first(x, c)], 1u);
    Softmax s;
    softmax_of(x, c, y, s);
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) {
      const float oh = c == 1 ? (float)y : (k == y ? 1.f : 0.f);   // C == 1 quirk, see head_loss.cu
      if (k < c) {
        P[k] += s.p[k];
        I[k] = fmaf(s.p[k], oh, I[k]);
        G[k] += oh;
      }
    }
    if (!(a.has_ignore && yl == a.ignore_index)) {
      const float wy = a.class_weights ? a.class_weights[y] : 1.f;
      cen = fmaf(-s.logp_y, wy, cen);
      ced += wy;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) {
    const float i = warp_sum(I[k]), pp = warp_sum(P[k]), g = warp_sum(G[k]);
    if (lane == 0) {
      red[warp][k] = i;
      red[warp][kMaxClasses + k] = pp;
      red[warp][2 * kMaxClasses + k] = g;
    }
  }
  cen = warp_sum(cen);
  ced = warp_sum(ced);
  if (lane == 0) {
    red[warp][3 * kMaxClasses] = cen;
    red[warp][3 * kMaxClasses + 1] = ced;
  }
  __syncthreads();
  double* accum = a.accum + (size_t)blockIdx.y * (3 * c + 2);
  if (threadIdx.x < 3 * kMaxClasses + 2) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    const int q = threadIdx.x / kMaxClasses, k = threadIdx.x % kMaxClasses;
    if (threadIdx.x >= 3 * kMaxClasses)
      atomicAdd(accum + 3 * c + (threadIdx.x - 3 * kMaxClasses), (double)s);
    else if (k < c)
      atomicAdd(accum + q * c + k, (double)s);
  }
  if (threadIdx.x < c) {
    const int k = threadIdx.x;
    unsigned long long tp = 0, fp = 0, fn = 0, total = 0;
    for (int l = 0; l < c; ++l)
      for (int q = 0; q < c; ++q) {
        const unsigned long long v = cm[l * kMaxClasses + q];
        total += v;
        if (l == k && q == k) tp += v;
        else if (q == k) fp += v;
        else if (l == k) fn += v;
      }
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(a.counts);
    const unsigned long long tn = total - tp - fp - fn;
    if (tp) atomicAdd(counts + 0 * c + k, tp);
    if (fp) atomicAdd(counts + 1 * c + k, fp);
    if (fn) atomicAdd(counts + 2 * c + k, fn);
    if (tn) atomicAdd(counts + 3 * c + k, tn);
  }
}

// per-image loss (utils/weighted_loss.py:76-98,165), then total_loss += loss.item() in image order (utils/training.py:98)
__global__ void eval_finalize_kernel(unetk_eval_args a) {
  for (int i = threadIdx.x; i < a.n; i += blockDim.x)
    a.loss_per_image[i] = finalize_dice_ce(a.accum + (size_t)i * (3 * a.c + 2), a.c, a.class_weights, a.has_ignore,
                                           a.ignore_index, a.smooth, a.dice_weight, a.ce_weight, nullptr);
  __syncthreads();
  if (threadIdx.x == 0 && a.loss_sum) {
    double tot = a.loss_sum[0];
    for (int i = 0; i < a.n; ++i) tot += (double)a.loss_per_image[i];
    a.loss_sum[0] = tot;
  }
}

static unsigned blocks_for(int max_out_pixels, int n) {
  int64_t bx = ((int64_t)max_out_pixels + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = ((int64_t)sm_count() * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  return (unsigned)(bx > 0 ? bx : 1);
}

}  // namespace unetk

using namespace unetk;

extern "C" {

int unetk_crop_resize(const float* src, int32_t n, int32_t c, int32_t th, int32_t tw, const unetk_eval_image* images,
                      int32_t max_out_pixels, int32_t mode, float* out_packed, void* stream) {
  UNETK_REQUIRE(src && images && out_packed, "crop_resize: null argument");
  UNETK_REQUIRE(n > 0 && n <= 65535 && th > 0 && tw > 0 && c >= 1 && c <= kMaxClasses && max_out_pixels > 0,
                "crop_resize: 1..8 channels, 1..65535 images");
  UNETK_REQUIRE(mode == 0 || mode == 1, "crop_resize: mode must be 0 (bilinear) or 1 (nearest)");
  const dim3 grid(blocks_for(max_out_pixels, n), (unsigned)n);
  if (mode == 0)
    crop_resize_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(src, c, th, tw, images, out_packed);
  else
    crop_resize_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(src, c, th, tw, images, out_packed);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_eval_loss_metrics(const unetk_eval_args* a, void* stream) {
  UNETK_REQUIRE(a && a->logits && a->images && a->labels && a->accum && a->loss_per_image && a->counts && a->status,
                "eval_loss_metrics: null argument");
  UNETK_REQUIRE(a->n > 0 && a->n <= 65535 && a->th > 0 && a->tw > 0 && a->c >= 1 && a->c <= kMaxClasses &&
                    a->max_out_pixels > 0, "eval_loss_metrics: 1..8 classes, 1..65535 images");
  UNETK_REQUIRE(a->label_dtype == UNETK_U8 || a->label_dtype == UNETK_I64, "eval_loss_metrics: labels must be u8 or i64");
  const dim3 grid(blocks_for(a->max_out_pixels, a->n), (unsigned)a->n);
  if (a->label_dtype == UNETK_U8)
    eval_reduce_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  else
    eval_reduce_kernel<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  UNETK_LAUNCH_CHECK();
  eval_finalize_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(*a);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}
}
