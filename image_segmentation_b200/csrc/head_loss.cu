// 1x1 classifier head, fused weighted Dice + cross-entropy loss, and argmax/confusion-count metrics.
//
// Reference call sites replaced:
//   unet/unet.py:91                      nn.Conv2d(64, dout, 1) forward/backward
//   utils/weighted_loss.py:36-98,163     softmax, scatter_ one-hot, Dice sums, clip, weighted mean, cross_entropy
//   utils/MetricsHistory.py:65-86        argmax, one_hot x2, four masked sums, four .cpu() copies per image
//
// Roofline: HBM.  Algorithmic bytes per pixel:
//   head fprop   cin*sizeof(T) + 4*dout                 head bwd   2*cin*sizeof(T) + 4*dout (a read once, da written once)
//   loss fwd     4*C + 8                                loss bwd   8*C + 8
//   metrics      4*C + 8
#include "loss_common.cuh"

namespace unetk {

constexpr int kMaxHeadCin = 256;

// ------------------------------------------------------------------------------------------------
// head
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) head_fprop_kernel(const T* __restrict__ a, int ld, int cin, int64_t npix,
                                                         int64_t hw, const float* __restrict__ w,
                                                         const float* __restrict__ b, int dout,
                                                         float* __restrict__ logits) {
  __shared__ float ws[kMaxClasses * kMaxHeadCin];
  __shared__ float bs[kMaxClasses];
  for (int i = threadIdx.x; i < dout * cin; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < dout) bs[threadIdx.x] = b ? b[threadIdx.x] : 0.f;
  __syncthreads();
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    float acc[kMaxClasses];
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) acc[k] = k < dout ? bs[k] : 0.f;
    for (int g = 0; g < cin / 8; ++g) {
      float v[8];
      load8(a + p * ld + g * 8, v);
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k) {
        if (k < dout) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k] = fmaf(v[j], ws[k * cin + g * 8 + j], acc[k]);
        }
      }
    }
    const int64_t img = p / hw, off = p % hw;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < dout) logits[(img * dout + k) * hw + off] = acc[k];
  }
}

// Fused head backward.  One block walks tiles of TILE pixels:
//   (1) thread-per-pixel: read dl[p, :] (coalesced per class plane), write da[p, :] = dl . W   (128 B per thread)
//   (2) cooperative, fully coalesced copy of the activation tile a[TILE][cin] into shared memory
//   (3) thread (k, ci): dw[k, ci] += sum_p dl[p, k] * a[p, ci]  from shared memory (conflict-free)
// db[k] is reduced from the per-thread dl values.  dw/db are flushed with one atomicAdd per block.
template <typename T, int TILE>
__global__ void __launch_bounds__(TILE) head_bwd_kernel(const float* __restrict__ dl, const T* __restrict__ a, int ald,
                                                       int cin, int64_t npix, int64_t hw, const float* __restrict__ w,
                                                       int dout, T* __restrict__ da, int dald, float* __restrict__ dw,
                                                       float* __restrict__ db) {
  extern __shared__ __align__(16) uint8_t head_smem[];
  T* a_s = reinterpret_cast<T*>(head_smem);                                   // [TILE][cin]
  float* dl_s = reinterpret_cast<float*>(head_smem + (size_t)TILE * cin * sizeof(T));  // [kMaxClasses][TILE]
  float* ws = dl_s + kMaxClasses * TILE;                                     // [dout][cin]
  for (int i = threadIdx.x; i < dout * cin; i += TILE) ws[i] = w[i];
  __syncthreads();
  const int wk = threadIdx.x / cin, wci = threadIdx.x % cin;                 // role in step (3)
  const bool w_active = threadIdx.x < dout * cin;
  float wacc = 0.f;
  float wacc2 = 0.f;   // second (k, ci) pair when dout*cin > TILE
  float bsum[kMaxClasses];
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) bsum[k] = 0.f;
  const int64_t ntiles = (npix + TILE - 1) / TILE;
  const bool contiguous = ald == cin;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * TILE;
    const int64_t p = p0 + threadIdx.x;
    const bool ok = p < npix;
    float g[kMaxClasses];
    {
      const int64_t img = ok ? p / hw : 0, off = ok ? p % hw : 0;
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k) {
        g[k] = (ok && k < dout) ? dl[(img * dout + k) * hw + off] : 0.f;
        bsum[k] += g[k];
        dl_s[k * TILE + threadIdx.x] = g[k];
      }
    }
    // (2) stage the activation tile
    if (contiguous) {
      const int64_t tile_elems = min((int64_t)TILE, npix - p0) * cin;
      const int vec = 16 / (int)sizeof(T);
      const T* src = a + p0 * ald;
      for (int64_t i = (int64_t)threadIdx.x * vec; i < (int64_t)TILE * cin; i += (int64_t)TILE * vec) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (i < tile_elems) v = *reinterpret_cast<const uint4*>(src + i);
        *reinterpret_cast<uint4*>(a_s + i) = v;
      }
    } else {
      for (int c0 = 0; c0 < cin; c0 += 8) {
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (ok) load8(a + p * ald + c0, v);
        store8(a_s + (size_t)threadIdx.x * cin + c0, v);
      }
    }
    // (1) data gradient
    if (ok) {
      for (int gidx = 0; gidx < cin / 8; ++gidx) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) {
          if (k < dout) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(g[k], ws[k * cin + gidx * 8 + j], o[j]);
          }
        }
        store8(da + p * dald + gidx * 8, o);
      }
    }
    __syncthreads();
    // (3) weight gradient from shared memory
    if (w_active) {
      float s0 = 0.f;
#pragma unroll 8
      for (int q = 0; q < TILE; ++q) s0 = fmaf(dl_s[wk * TILE + q], to_f(a_s[(size_t)q * cin + wci]), s0);
      wacc += s0;
    }
    if (threadIdx.x + TILE < dout * cin) {
      const int k2 = (threadIdx.x + TILE) / cin, c2 = (threadIdx.x + TILE) % cin;
      float s0 = 0.f;
#pragma unroll 8
      for (int q = 0; q < TILE; ++q) s0 = fmaf(dl_s[k2 * TILE + q], to_f(a_s[(size_t)q * cin + c2]), s0);
      wacc2 += s0;
    }
    __syncthreads();
  }
  if (w_active) atomicAdd(dw + wk * cin + wci, wacc);
  if (threadIdx.x + TILE < dout * cin) atomicAdd(dw + threadIdx.x + TILE, wacc2);
  if (db) {
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) {
      if (k < dout) {
        const float sres = warp_sum(bsum[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(db + k, sres);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Dice + CE
// ------------------------------------------------------------------------------------------------
template <int CM>
__global__ void __launch_bounds__(256) dice_ce_reduce_kernel(unetk_dice_ce_args a) {
  const int c = a.c;
  const int64_t hw = (int64_t)a.h * a.w, npix = (int64_t)a.n * hw;
  float I[CM], P[CM], G[CM], cen = 0.f, ced = 0.f;
#pragma unroll
  for (int k = 0; k < CM; ++k) I[k] = P[k] = G[k] = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t yl = a.target[p];
    if (yl < 0 || yl >= c) {
      atomicOr(a.status, 1);
      continue;
    }
    const int y = (int)yl;
    const int64_t img = p / hw, off = p % hw;
    SoftmaxT<CM> s;
    if (a.input_kind == UNETK_LOSS_LOGITS)
      pixel_softmax<CM>(a.logits, img * c * hw + off, hw, c, y, s);
    else
      pixel_probs<CM>(a.logits, img * c * hw + off, hw, c, y, a.input_kind, a.nll_eps, s);
#pragma unroll
    for (int k = 0; k < CM; ++k) {
      // utils/weighted_loss.py:50-58: with C == 1 the reference uses y.float() itself as the "one-hot"
      const float oh = c == 1 ? (float)y : (k == y ? 1.f : 0.f);
      if (k < c) {
        P[k] += s.p[k];
        I[k] = fmaf(s.p[k], oh, I[k]);
        G[k] += oh;
      }
    }
    const bool valid = !(a.has_ignore && yl == a.ignore_index);
    if (valid) {
      const float wy = a.class_weights ? a.class_weights[y] : 1.f;
      cen = fmaf(-s.logp_y, wy, cen);
      ced += wy;
    }
  }
  __shared__ float red[8][3 * CM + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < CM; ++k) {
    const float i = warp_sum(I[k]), pp = warp_sum(P[k]), g = warp_sum(G[k]);
    if (lane == 0) {
      red[warp][k] = i;
      red[warp][CM + k] = pp;
      red[warp][2 * CM + k] = g;
    }
  }
  cen = warp_sum(cen);
  ced = warp_sum(ced);
  if (lane == 0) {
    red[warp][3 * CM] = cen;
    red[warp][3 * CM + 1] = ced;
  }
  __syncthreads();
  if (threadIdx.x < 3 * CM + 2) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    const int q = threadIdx.x / CM, k = threadIdx.x % CM;
    if (threadIdx.x >= 3 * CM)
      atomicAdd(a.accum + 3 * c + (threadIdx.x - 3 * CM), (double)s);
    else if (k < c)
      atomicAdd(a.accum + q * c + k, (double)s);
  }
}

__global__ void dice_ce_finalize_kernel(unetk_dice_ce_args a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  a.loss[0] = finalize_dice_ce(a.accum, a.c, a.class_weights, a.has_ignore, a.ignore_index, a.smooth, a.dice_weight,
                               a.ce_weight, a.coef);
}

template <int CM>
__global__ void __launch_bounds__(256) dice_ce_bwd_kernel(unetk_dice_ce_args a) {
  const int c = a.c;
  const int64_t hw = (int64_t)a.h * a.w, npix = (int64_t)a.n * hw;
  __shared__ float coef[2 * CM + 1];
  if (threadIdx.x < 2 * c + 1) coef[threadIdx.x] = a.coef[threadIdx.x];
  __syncthreads();
  const float go = a.grad_out ? a.grad_out[0] : 1.f;
  const float ce_scale = coef[2 * c];
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t yl = a.target[p];
    const int64_t img = p / hw, off = p % hw;
    const int64_t base = img * c * hw + off;
    if (yl < 0 || yl >= c) {
      for (int k = 0; k < c; ++k) a.dlogits[base + k * hw] = 0.f;
      continue;
    }
    const int y = (int)yl;
    SoftmaxT<CM> s;
    if (a.input_kind != UNETK_LOSS_LOGITS) {
      // probability inputs (utils/weighted_loss.py:207-210,338-340): no softmax Jacobian;
      //   d Dice / d p_k = -coef_k (2 oh_k - dc_k),   d NLL / d p_y = -w_y / (p_y + eps)   (or -w_y for raw inputs)
      pixel_probs<CM>(a.logits, base, hw, c, y, a.input_kind, a.nll_eps, s);
      const bool valid = !(a.has_ignore && yl == a.ignore_index);
      const float wy = valid ? (a.class_weights ? a.class_weights[y] : 1.f) * ce_scale : 0.f;
#pragma unroll
      for (int k = 0; k < CM; ++k) {
        if (k < c) {
          const float oh = c == 1 ? (float)y : (k == y ? 1.f : 0.f);
          float gx = -coef[k] * (2.f * oh - coef[c + k]);
          if (k == y) gx -= a.input_kind == UNETK_LOSS_PROBS_LOG ? wy / (s.p[k] + a.nll_eps) : wy;
          a.dlogits[base + k * hw] = gx * go;
        }
      }
      continue;
    }
    pixel_softmax<CM>(a.logits, base, hw, c, y, s);
    float g[CM], dot = 0.f;
#pragma unroll
    for (int k = 0; k < CM; ++k) {
      const float oh = c == 1 ? (float)y : (k == y ? 1.f : 0.f);
      g[k] = k < c ? -coef[k] * (2.f * oh - coef[c + k]) : 0.f;
      dot = fmaf(s.p[k], g[k], dot);
    }
    const bool valid = !(a.has_ignore && yl == a.ignore_index);
    const float wy = valid ? (a.class_weights ? a.class_weights[y] : 1.f) * ce_scale : 0.f;
#pragma unroll
    for (int k = 0; k < CM; ++k) {
      if (k < c) {
        const float gx = s.p[k] * (g[k] - dot) + wy * (s.p[k] - (k == y ? 1.f : 0.f));
        a.dlogits[base + k * hw] = gx * go;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// metrics
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    argmax_confusion_kernel(const float* __restrict__ pred, const int64_t* __restrict__ label, int n, int c, int64_t hw,
                            unsigned long long* __restrict__ counts, uint8_t* __restrict__ argmax_out,
                            int* __restrict__ status) {
  __shared__ unsigned int cm[kMaxClasses * kMaxClasses];
  for (int i = threadIdx.x; i < kMaxClasses * kMaxClasses; i += blockDim.x) cm[i] = 0;
  __syncthreads();
  const int64_t npix = (int64_t)n * hw;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = p / hw, off = p % hw;
    const float* px = pred + img * c * hw + off;
    float best = px[0];
    int arg = 0;
    for (int k = 1; k < c; ++k) {
      const float v = px[k * hw];
      // torch.argmax: first maximum wins, NaN counts as the largest value
      if (v > best || (isnan(v) && !isnan(best))) {
        best = v;
        arg = k;
      }
    }
    if (argmax_out) argmax_out[p] = (uint8_t)arg;
    const int64_t yl = label[p];
    if (yl < 0 || yl >= c) {
      atomicOr(status, 1);
      continue;
    }
    atomicAdd(&cm[(int)yl * kMaxClasses + arg], 1u);
  }
  __syncthreads();
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
    const unsigned long long tn = total - tp - fp - fn;
    if (tp) atomicAdd(counts + 0 * c + k, tp);
    if (fp) atomicAdd(counts + 1 * c + k, fp);
    if (fn) atomicAdd(counts + 2 * c + k, fn);
    if (tn) atomicAdd(counts + 3 * c + k, tn);
  }
}

static int grid_pixels(int64_t npix, int per_thread) {
  int64_t blocks = (npix + 256LL * per_thread - 1) / (256LL * per_thread);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks > 0 ? blocks : 1);
}

}  // namespace unetk

using namespace unetk;

extern "C" {

int unetk_head_fprop(const unetk_tensor* a, const float* w, const float* b, int32_t dout, float* logits_nchw,
                     void* stream) {
  UNETK_REQUIRE(a && w && logits_nchw, "head_fprop: null argument");
  UNETK_REQUIRE(tensor_ok(*a) && vec8_ok(*a), "head_fprop: a must be NHWC with c%%8==0");
  UNETK_REQUIRE(dout >= 1 && dout <= kMaxClasses && a->c <= kMaxHeadCin, "head_fprop: dout<=8, cin<=256 supported");
  const int64_t npix = pixels(*a), hw = (int64_t)a->h * a->w;
  UNETK_DISPATCH_DTYPE(a->dtype, T, {
    head_fprop_kernel<T><<<grid_pixels(npix, 1), 256, 0, (cudaStream_t)stream>>>((const T*)a->ptr, a->ld, a->c, npix, hw, w, b, dout, logits_nchw);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_head_bwd(const float* dlogits_nchw, const unetk_tensor* a, const float* w, int32_t dout, const unetk_tensor* da,
                   float* dw, float* db, void* stream) {
  UNETK_REQUIRE(dlogits_nchw && a && w && da && dw, "head_bwd: null argument");
  UNETK_REQUIRE(tensor_ok(*a) && vec8_ok(*a) && tensor_ok(*da) && vec8_ok(*da), "head_bwd: bad tensor");
  UNETK_REQUIRE(da->dtype == a->dtype && da->n == a->n && da->h == a->h && da->w == a->w && da->c == a->c,
                "head_bwd: da must match a");
  UNETK_REQUIRE(dout >= 1 && dout <= kMaxClasses && a->c <= 128, "head_bwd: dout<=8 and cin<=128 supported");
  const int64_t npix = pixels(*a), hw = (int64_t)a->h * a->w;
  constexpr int TILE = 256;
  UNETK_REQUIRE(dout * a->c <= 2 * TILE, "head_bwd: dout*cin must be <= 512");
  const size_t es = a->dtype == UNETK_BF16 ? 2 : 4;
  const size_t smem = (size_t)TILE * a->c * es + (size_t)kMaxClasses * TILE * 4 + (size_t)dout * a->c * 4;
  UNETK_REQUIRE(smem <= 100 * 1024, "head_bwd: activation tile does not fit shared memory (cin * element size too large)");
  const int64_t ntiles = (npix + TILE - 1) / TILE;
  int64_t grid = ntiles < (int64_t)sm_count() * 3 ? ntiles : (int64_t)sm_count() * 3;
  if (grid < 1) grid = 1;
  UNETK_DISPATCH_DTYPE(a->dtype, T, {
    static bool attr_set = false;
    if (!attr_set) {
      UNETK_CUDA(cudaFuncSetAttribute(head_bwd_kernel<T, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set = true;
    }
    head_bwd_kernel<T, TILE><<<(unsigned)grid, TILE, smem, (cudaStream_t)stream>>>(
        dlogits_nchw, (const T*)a->ptr, a->ld, a->c, npix, hw, w, dout, (T*)da->ptr, da->ld, dw, db);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

static int check_loss(const unetk_dice_ce_args* a) {
  UNETK_REQUIRE(a && a->logits && a->target && a->coef && a->status, "dice_ce: null argument");
  UNETK_REQUIRE(a->n > 0 && a->h > 0 && a->w > 0 && a->c >= 1 && a->c <= kMaxClasses, "dice_ce: 1..8 classes supported");
  return UNETK_OK;
}

int unetk_dice_ce_fwd(const unetk_dice_ce_args* a, void* stream) {
  int rc = check_loss(a);
  if (rc) return rc;
  UNETK_REQUIRE(a->accum && a->loss, "dice_ce_fwd: null accum/loss");
  const int64_t npix = (int64_t)a->n * a->h * a->w;
  if (a->c <= 4)
    dice_ce_reduce_kernel<4><<<grid_pixels(npix, 4), 256, 0, (cudaStream_t)stream>>>(*a);
  else
    dice_ce_reduce_kernel<kMaxClasses><<<grid_pixels(npix, 4), 256, 0, (cudaStream_t)stream>>>(*a);
  dice_ce_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_dice_ce_bwd(const unetk_dice_ce_args* a, void* stream) {
  int rc = check_loss(a);
  if (rc) return rc;
  UNETK_REQUIRE(a->dlogits, "dice_ce_bwd: null dlogits");
  const int64_t npix = (int64_t)a->n * a->h * a->w;
  if (a->c <= 4)
    dice_ce_bwd_kernel<4><<<grid_pixels(npix, 2), 256, 0, (cudaStream_t)stream>>>(*a);
  else
    dice_ce_bwd_kernel<kMaxClasses><<<grid_pixels(npix, 2), 256, 0, (cudaStream_t)stream>>>(*a);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_argmax_confusion(const float* pred, const int64_t* label, int32_t n, int32_t c, int32_t h, int32_t w,
                           int64_t* counts, uint8_t* argmax_out, int32_t* status, void* stream) {
  UNETK_REQUIRE(pred && label && counts && status, "argmax_confusion: null argument");
  UNETK_REQUIRE(n > 0 && h > 0 && w > 0 && c >= 1 && c <= kMaxClasses, "argmax_confusion: 1..8 classes supported");
  const int64_t hw = (int64_t)h * w;
  argmax_confusion_kernel<<<grid_pixels((int64_t)n * hw, 4), 256, 0, (cudaStream_t)stream>>>(
      pred, label, n, c, hw, reinterpret_cast<unsigned long long*>(counts), argmax_out, status);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}
}
