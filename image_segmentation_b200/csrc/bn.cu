// BatchNorm(train) + ReLU + MaxPool2x2 forward/backward as bandwidth kernels (NHWC, 8 channels per thread).
//
// Reference call sites replaced: unet/unet.py:17-18,20-21 (BatchNorm2d + ReLU inside DoubleConvReLU) and
// unet/unet.py:40 (MaxPool2d(2,2) in Down) together with their autograd backward
// (native_batch_norm_backward, threshold_backward, max_pool2d_with_indices_backward).
//
// Roofline: HBM.  Algorithmic bytes per element (T = storage type):
//   stats            read z                                  1*sizeof(T)
//   apply(+pool)     read z, write a (+ pooled/4)            2*sizeof(T) (+ sizeof(T)/4)
//   bwd reduce       read z, dy (+ dpool/4)                  2*sizeof(T) (+ sizeof(T)/4)
//   bwd apply        read z, dy (+ dpool/4), write dz        3*sizeof(T) (+ sizeof(T)/4)
#include "common.cuh"

namespace unetk {

constexpr int kThreads = 256;

// largest power of two <= 32 dividing cg
static int choose_cgb(int cg) {
  int b = 32;
  while (b > 1 && (cg % b) != 0) b >>= 1;
  return b;
}

// ------------------------------------------------------------------------------------------------
// per-channel reductions: block = rows x CGB channel-groups; smem reduce across rows; double atomics
// ------------------------------------------------------------------------------------------------
template <int NQ>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NQ][8], int cgb, int row, int lane_g, int rows,
                                                     int cg0, double* const (&out)[NQ]) {
  extern __shared__ float red[];  // [NQ][rows][cgb*8]
  const int width = cgb * 8;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(q * rows + row) * width + lane_g * 8 + i] = acc[q][i];
  __syncthreads();
  for (int idx = threadIdx.x; idx < NQ * width; idx += blockDim.x) {
    const int q = idx / width, ch = idx % width;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += red[(q * rows + r) * width + ch];
    atomicAdd(out[q] + (size_t)cg0 * 8 + ch, (double)s);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const T* __restrict__ z, int64_t npix, int ld, int cgb,
                                                            int pix_per_block, double* __restrict__ sum,
                                                            double* __restrict__ sumsq) {
  const int rows = kThreads / cgb;
  const int lane_g = threadIdx.x % cgb, row = threadIdx.x / cgb;
  const int cg0 = blockIdx.x * cgb;
  const int64_t p0 = (int64_t)blockIdx.y * pix_per_block;
  const int64_t p1 = min(p0 + (int64_t)pix_per_block, npix);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  for (int64_t p = p0 + row; p < p1; p += rows) {
    float v[8];
    load8(z + p * ld + (size_t)(cg0 + lane_g) * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[0][i] += v[i];
      acc[1][i] = fmaf(v[i], v[i], acc[1][i]);
    }
  }
  double* const outs[2] = {sum, sumsq};
  block_channel_reduce<2>(acc, cgb, row, lane_g, rows, cg0, outs);
}

// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(unetk_bn_finalize_args a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && a.training && a.num_batches_tracked) *a.num_batches_tracked += 1;
  if (c >= a.c) return;
  const float g = a.gamma[c], b = a.beta[c];
  const float cb = a.conv_bias ? a.conv_bias[c] : 0.f;
  if (a.training) {
    const double m = a.sum[c] / (double)a.count;
    double var = a.sumsq[c] / (double)a.count - m * m;
    if (var < 0.0) var = 0.0;
    const double istd = 1.0 / sqrt(var + (double)a.eps);
    const float sc = (float)((double)g * istd);
    a.scale[c] = sc;
    a.shift[c] = (float)((double)b - m * (double)g * istd);
    a.mean[c] = (float)m;
    a.invstd[c] = (float)istd;
    if (a.running_mean) {
      // PyTorch: running = (1-momentum)*running + momentum*batch; variance uses the unbiased estimate
      const double unb = a.count > 1 ? var * ((double)a.count / (double)(a.count - 1)) : var;
      a.running_mean[c] = (float)((1.0 - a.momentum) * a.running_mean[c] + a.momentum * (m + cb));
      a.running_var[c] = (float)((1.0 - a.momentum) * a.running_var[c] + a.momentum * unb);
    }
  } else {
    const float istd = rsqrtf(a.running_var[c] + a.eps);
    const float sc = g * istd;
    a.scale[c] = sc;
    a.shift[c] = fmaf(cb - a.running_mean[c], sc, b);
    if (a.mean) a.mean[c] = a.running_mean[c] - cb;
    if (a.invstd) a.invstd[c] = istd;
  }
}

// Two-level index (img, off) of a linear item index over rows of `hw` items, advanced by a fixed stride from round to round
// of a loop without a division: (image, pixel offset) of a warp's 16-pixel block inside NCHW planes in the head kernels
// (the divisions by the runtime plane size were a third of their loops), (image row, pixel pair) / (pooled row, window) in
// the pooling variants of the BatchNorm kernels.
struct BlockCursor {
  uint32_t img, off, step_img, step_off, hw;
  __device__ __forceinline__ BlockCursor(int64_t base, int64_t stride, uint32_t hw_) : hw(hw_) {
    img = (uint32_t)(base / hw_);
    off = (uint32_t)(base - (int64_t)img * hw_);
    step_img = (uint32_t)(stride / hw_);
    step_off = (uint32_t)(stride - (int64_t)step_img * hw_);
  }
  __device__ __forceinline__ void advance() {
    img += step_img;
    off += step_off;
    if (off >= hw) {
      off -= hw;
      ++img;
    }
  }
  // pixel (block base + t), t < 16
  __device__ __forceinline__ void at(uint32_t t, uint32_t& i, uint32_t& o) const {
    i = img;
    o = off + t;
    while (o >= hw) {
      o -= hw;
      ++i;
    }
  }
};

// ------------------------------------------------------------------------------------------------
template <typename T, bool POOL>
__global__ void __launch_bounds__(kThreads)
    bn_relu_apply_kernel(const T* __restrict__ z, int zld, const float* __restrict__ scale,
                         const float* __restrict__ shift, T* __restrict__ a, int ald, T* __restrict__ pooled, int pld,
                         uint16_t* __restrict__ pool_idx, int n, int h, int w, int cg) {
  // work item = (pixel or 2x2 window, channel group)
  const int hh = POOL ? h / 2 : h, ww = POOL ? w / 2 : w;
  const int64_t total = (int64_t)n * hh * ww * cg;
  // POOL with cg | blockDim (every power-of-two channel count): the thread's channel group is fixed and its window index
  // advances by a constant, so (pooled row, window) come from a cursor instead of four divisions per item
  const bool fixed_g = POOL && (blockDim.x % cg) == 0;
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  BlockCursor cur(fixed_g ? first / cg : 0, fixed_g ? ((int64_t)gridDim.x * blockDim.x) / cg : 1, (uint32_t)ww);
  const int g_fixed = (int)(first % cg);
  for (int64_t i = first; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = fixed_g ? g_fixed : (int)((uint32_t)i % (uint32_t)cg);
    int64_t p = fixed_g ? 0 : (int64_t)((uint32_t)i / (uint32_t)cg);
    float sc[8], sh[8];
    load8(scale + g * 8, sc);
    load8(shift + g * 8, sh);
    if (!POOL) {
      float v[8];
      load8(z + p * zld + g * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f);
      store8(a + p * ald + g * 8, v);
    } else {
      // line = img * hh + y (pooled row); the full-resolution rows are 2 * line and 2 * line + 1 because h == 2 * hh
      uint32_t line, x;
      if (fixed_g) {
        line = cur.img;
        x = cur.off;
        cur.advance();
      } else {
        const uint32_t p32 = (uint32_t)p;
        line = p32 / (uint32_t)ww;
        x = p32 - line * (uint32_t)ww;
      }
      float mx[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) mx[k] = -1.f;  // ReLU output is >= 0: position 0 always wins the first comparison
      uint32_t arg = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t pix = ((int64_t)(2 * line + (q >> 1))) * w + (2 * x + (q & 1));
        float v[8];
        load8(z + pix * zld + g * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = round_to<T>(fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f));
          if (v[k] > mx[k]) {   // strict: the first maximum in scan order keeps the gradient (torch's max_pool2d rule)
            mx[k] = v[k];
            arg = (arg & ~(3u << (2 * k))) | ((uint32_t)q << (2 * k));
          }
        }
        store8(a + pix * ald + g * 8, v);
      }
      const int64_t win = (int64_t)line * ww + x;
      store8(pooled + win * pld + g * 8, mx);
      if (pool_idx) pool_idx[win * cg + g] = (uint16_t)arg;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: dy = (dA_full + routed pool grad) * [a > 0]
// ------------------------------------------------------------------------------------------------
// [a > 0] for the activation the forward pass STORED, a = T(max(z*sc + sh, 0)), decided without redoing the rounding:
// a float v rounds to a positive T exactly when v > threshold (bf16: values up to 2^-134 round to zero, ties to even).
template <typename T> __device__ __forceinline__ float act_threshold() { return 0.f; }
template <> __device__ __forceinline__ float act_threshold<__nv_bfloat16>() { return 0x1p-134f; }

template <typename T>
struct BwdSrc {
  const T* z; int zld;
  const T* dy; int dyld;      // may be null
  const T* dp; int dpld;      // may be null (POOL variants only)
  const uint16_t* pidx;       // [N,H/2,W/2,C/8]: 2-bit window position per channel, written by bn_relu_apply (POOL only)
  const float* scale; const float* shift; const float* mean; const float* invstd;
  int n, h, w, cg;
};

// masked upstream gradient of one pixel, 8 channels:  dy = (dA_full + routed pooled gradient) * [a > 0]
// POOL: the pooled gradient of window (y/2, x/2) goes to the position recorded by the forward pass (pool_idx: 2 bits per
// channel, first maximum in scan order), so the backward pass never re-derives the argmax.
template <typename T, bool POOL>
__device__ __forceinline__ void load_pixel(const BwdSrc<T>& s, int64_t pix, int g, const float (&sc)[8], const float (&sh)[8],
                                           float (&zv)[8], float (&dy)[8]) {
  load8(s.z + pix * s.zld + g * 8, zv);
  float d[8];
  if (!POOL || s.dy) {
    load8(s.dy + pix * s.dyld + g * 8, d);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = 0.f;
  }
  if (POOL) {
    // 32-bit index arithmetic (the host checks pixels < 2^31): 64-bit div/mod made this path compute bound
    const uint32_t p32 = (uint32_t)pix, uw = (uint32_t)s.w, uh = (uint32_t)s.h;
    const uint32_t line = p32 / uw, x = p32 - line * uw;
    const uint32_t img = line / uh, y = line - img * uh;
    const int64_t win = ((int64_t)(img * (uh >> 1) + (y >> 1))) * (uw >> 1) + (x >> 1);
    const uint32_t q = (y & 1u) * 2u + (x & 1u);
    const uint32_t arg = s.pidx[win * s.cg + g];
    float dp[8];
    load8(s.dp + win * s.dpld + g * 8, dp);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] += (((arg >> (2 * k)) & 3u) == q) ? dp[k] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) dy[k] = fmaf(zv[k], sc[k], sh[k]) > act_threshold<T>() ? d[k] : 0.f;
}

// POOL: two horizontally adjacent pixels (2*xp, 2*xp+1) of one pooling window share the window's pooled gradient and
// arg-max word: one index decode, one dpool load and one idx load per pair.
template <typename T>
__device__ __forceinline__ void load_pixel_pair(const BwdSrc<T>& s, uint32_t line, uint32_t xp, int g, const float (&sc)[8],
                                                const float (&sh)[8], int64_t& pix0, float (&za)[8], float (&ya)[8],
                                                float (&zb)[8], float (&yb)[8]) {
  // line = img * h + y (image row), xp = pixel pair inside the row.  h is even, so the row's parity is y's and the pooled
  // row of the window is line / 2: no further decoding is needed.
  const uint32_t hw2 = (uint32_t)s.w >> 1;
  pix0 = (int64_t)line * s.w + 2 * xp;
  const int64_t win = (int64_t)(line >> 1) * hw2 + xp;
  const uint32_t q0 = (line & 1u) * 2u;
  const uint32_t arg = s.pidx[win * s.cg + g];
  float dp[8], da[8], db[8];
  load8(s.dp + win * s.dpld + g * 8, dp);
  load8(s.z + pix0 * s.zld + g * 8, za);
  load8(s.z + (pix0 + 1) * s.zld + g * 8, zb);
  if (s.dy) {
    load8(s.dy + pix0 * s.dyld + g * 8, da);
    load8(s.dy + (pix0 + 1) * s.dyld + g * 8, db);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) da[k] = db[k] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t a = (arg >> (2 * k)) & 3u;
    const float ga = da[k] + (a == q0 ? dp[k] : 0.f);
    const float gb = db[k] + (a == q0 + 1u ? dp[k] : 0.f);
    ya[k] = fmaf(za[k], sc[k], sh[k]) > act_threshold<T>() ? ga : 0.f;
    yb[k] = fmaf(zb[k], sc[k], sh[k]) > act_threshold<T>() ? gb : 0.f;
  }
}

template <typename T, bool POOL>
__global__ void __launch_bounds__(kThreads, 3)
    bn_bwd_reduce_kernel(BwdSrc<T> s, int cgb, int items_per_block, double* __restrict__ s1, double* __restrict__ s2) {
  const int rows = kThreads / cgb;
  const int lane_g = threadIdx.x % cgb, row = threadIdx.x / cgb;
  const int cg0 = blockIdx.x * cgb;
  const int g = cg0 + lane_g;
  const int64_t nitems = (int64_t)s.n * s.h * s.w / (POOL ? 2 : 1);
  const int64_t i0 = (int64_t)blockIdx.y * items_per_block;
  const int64_t i1 = min(i0 + (int64_t)items_per_block, nitems);
  float sc[8], sh[8];
  load8(s.scale + g * 8, sc);
  load8(s.shift + g * 8, sh);
  // accumulate sum(dy) and sum(dy*z); xhat = (z-mean)*invstd is applied once at the end:
  //   sum(dy*xhat) = invstd * (sum(dy*z) - mean*sum(dy))
  float acc[2][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[0][k] = acc[1][k] = 0.f;
  if (POOL) {
    // items are horizontal pixel pairs
    BlockCursor cur(i0 + row, rows, (uint32_t)s.w >> 1);   // (image row, pixel pair) of item `it`
    for (int64_t it = i0 + row; it < i1; it += rows, cur.advance()) {
      float za[8], ya[8], zb[8], yb[8];
      int64_t pix0;
      load_pixel_pair<T>(s, cur.img, cur.off, g, sc, sh, pix0, za, ya, zb, yb);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[0][k] += ya[k] + yb[k];
        acc[1][k] = fmaf(ya[k], za[k], fmaf(yb[k], zb[k], acc[1][k]));
      }
    }
  } else {
    int64_t it = i0 + row;
    for (; it + rows < i1; it += 2 * rows) {   // two pixels per iteration: more independent loads in flight
      float za[8], ya[8], zb[8], yb[8];
      load_pixel<T, false>(s, it, g, sc, sh, za, ya);
      load_pixel<T, false>(s, it + rows, g, sc, sh, zb, yb);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[0][k] += ya[k] + yb[k];
        acc[1][k] = fmaf(ya[k], za[k], fmaf(yb[k], zb[k], acc[1][k]));
      }
    }
    for (; it < i1; it += rows) {
      float zv[8], dy[8];
      load_pixel<T, false>(s, it, g, sc, sh, zv, dy);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[0][k] += dy[k];
        acc[1][k] = fmaf(dy[k], zv[k], acc[1][k]);
      }
    }
  }
  {
    float mu[8], is[8];
    load8(s.mean + g * 8, mu);
    load8(s.invstd + g * 8, is);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[1][k] = is[k] * (acc[1][k] - mu[k] * acc[0][k]);
  }
  double* const outs[2] = {s1, s2};
  block_channel_reduce<2>(acc, cgb, row, lane_g, rows, cg0, outs);
}

// Channel-stationary apply: a thread owns 8 channels for its whole pixel range, so the per-channel constants live in
// registers.  dz = sc*(dy - m1 - xhat*m2) is evaluated as  A*dy + B*z + C  with
//   A = sc,  B = -sc*m2*invstd,  C = sc*(m2*invstd*mean - m1).
template <typename T, bool POOL>
__global__ void __launch_bounds__(kThreads, 3)
    bn_bwd_apply_kernel(BwdSrc<T> s, int cgb, int items_per_block, const double* __restrict__ s1,
                        const double* __restrict__ s2, double inv_count, T* __restrict__ dz, int dzld,
                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int rows = kThreads / cgb;
  const int lane_g = threadIdx.x % cgb, row = threadIdx.x / cgb;
  const int g = blockIdx.x * cgb + lane_g;
  const int64_t nitems = (int64_t)s.n * s.h * s.w / (POOL ? 2 : 1);
  const int64_t i0 = (int64_t)blockIdx.y * items_per_block;
  const int64_t i1 = min(i0 + (int64_t)items_per_block, nitems);
  float sc[8], sh[8], cb[8], cc[8];
  load8(s.scale + g * 8, sc);
  load8(s.shift + g * 8, sh);
  {
    float mu[8], is[8];
    load8(s.mean + g * 8, mu);
    load8(s.invstd + g * 8, is);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m1 = (float)(s1[g * 8 + k] * inv_count);
      const float m2 = (float)(s2[g * 8 + k] * inv_count);
      cb[k] = -sc[k] * m2 * is[k];
      cc[k] = sc[k] * (m2 * is[k] * mu[k] - m1);
    }
  }
  if (blockIdx.y == 0 && row == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (dgamma) dgamma[g * 8 + k] = (float)s2[g * 8 + k];
      if (dbeta) dbeta[g * 8 + k] = (float)s1[g * 8 + k];
    }
  }
  if (POOL) {
    BlockCursor cur(i0 + row, rows, (uint32_t)s.w >> 1);   // (image row, pixel pair) of item `it`
    for (int64_t it = i0 + row; it < i1; it += rows, cur.advance()) {
      float za[8], ya[8], zb[8], yb[8], oa[8], ob[8];
      int64_t pix0;
      load_pixel_pair<T>(s, cur.img, cur.off, g, sc, sh, pix0, za, ya, zb, yb);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        oa[k] = fmaf(sc[k], ya[k], fmaf(cb[k], za[k], cc[k]));
        ob[k] = fmaf(sc[k], yb[k], fmaf(cb[k], zb[k], cc[k]));
      }
      store8(dz + pix0 * dzld + g * 8, oa);
      store8(dz + (pix0 + 1) * dzld + g * 8, ob);
    }
  } else {
    int64_t it = i0 + row;
    for (; it + rows < i1; it += 2 * rows) {
      float za[8], ya[8], zb[8], yb[8], oa[8], ob[8];
      load_pixel<T, false>(s, it, g, sc, sh, za, ya);
      load_pixel<T, false>(s, it + rows, g, sc, sh, zb, yb);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        oa[k] = fmaf(sc[k], ya[k], fmaf(cb[k], za[k], cc[k]));
        ob[k] = fmaf(sc[k], yb[k], fmaf(cb[k], zb[k], cc[k]));
      }
      store8(dz + it * dzld + g * 8, oa);
      store8(dz + (it + rows) * dzld + g * 8, ob);
    }
    for (; it < i1; it += rows) {
      float zv[8], dy[8], o[8];
      load_pixel<T, false>(s, it, g, sc, sh, zv, dy);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(sc[k], dy[k], fmaf(cb[k], zv[k], cc[k]));
      store8(dz + it * dzld + g * 8, o);
    }
  }
}

static int grid_for(int64_t total_threads) {
  const int64_t blocks = (total_threads + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

template <typename T>
static BwdSrc<T> make_src(const unetk_bn_bwd_args* a) {
  BwdSrc<T> s;
  s.z = (const T*)a->z.ptr; s.zld = a->z.ld;
  s.dy = (const T*)a->dy.ptr; s.dyld = a->dy.ld;
  s.dp = (const T*)a->dpool.ptr; s.dpld = a->dpool.ld;
  s.pidx = static_cast<const uint16_t*>(a->pool_idx);
  s.scale = a->scale; s.shift = a->shift; s.mean = a->mean; s.invstd = a->invstd;
  s.n = a->z.n; s.h = a->z.h; s.w = a->z.w; s.cg = a->z.c / 8;
  return s;
}

static int check_bwd(const unetk_bn_bwd_args* a) {
  UNETK_REQUIRE(a != nullptr, "bn_bwd: null args");
  UNETK_REQUIRE(tensor_ok(a->z) && vec8_ok(a->z), "bn_bwd: z must be NHWC with c%%8==0, ld%%8==0, 16B aligned");
  UNETK_REQUIRE(a->dy.ptr || a->dpool.ptr, "bn_bwd: need dy and/or dpool");
  if (a->dy.ptr) {
    UNETK_REQUIRE(tensor_ok(a->dy) && vec8_ok(a->dy) && a->dy.dtype == a->z.dtype && a->dy.n == a->z.n &&
                      a->dy.h == a->z.h && a->dy.w == a->z.w && a->dy.c == a->z.c, "bn_bwd: dy shape/dtype mismatch");
  }
  if (a->dpool.ptr) {
    UNETK_REQUIRE(tensor_ok(a->dpool) && vec8_ok(a->dpool) && a->dpool.dtype == a->z.dtype && a->dpool.n == a->z.n &&
                      a->dpool.h * 2 == a->z.h && a->dpool.w * 2 == a->z.w && a->dpool.c == a->z.c,
                  "bn_bwd: dpool must be [N,H/2,W/2,C]");
    UNETK_REQUIRE(a->pool_idx != nullptr, "bn_bwd: dpool needs the pool_idx written by unetk_bn_relu_apply");
  }
  UNETK_REQUIRE(a->scale && a->shift && a->mean && a->invstd && a->sums, "bn_bwd: null statistics");
  UNETK_REQUIRE(pixels(a->z) < (1LL << 31), "bn_bwd: more than 2^31 pixels");
  return UNETK_OK;
}

// ------------------------------------------------------------------------------------------------
// Classifier head backward fused with the BatchNorm backward of the layer that feeds it.
// The head is a 1x1 conv  logits[p][k] = sum_c a[p][c] * Wh[k][c] + bh[k]  with  a = relu(z*scale + shift)  (the stored
// activation of the last block).  Given dlogits (NCHW fp32) both passes recompute, per pixel and channel,
//   da = sum_k dlogits[k] * Wh[k][c],   dy = da * [a > 0]
// so the gradient of the 64-channel activation is never written to or read from HBM:
//   reduce : reads z + dlogits     -> sum dy, sum dy*xhat, dWh[k][c] = sum_p dlogits[k]*a[c], dbh[k] = sum_p dlogits[k]
//   apply  : reads z + dlogits     -> writes dz (and dgamma, dbeta, dWh, dbh from the fp64 sums)
// versus head_bwd (read a, write da) + bn_bwd_reduce (read z, da) + bn_bwd_apply (read z, da, write dz): 3.2 instead of
// 7 passes over a [N,H,W,64] tensor.
// ------------------------------------------------------------------------------------------------
static_assert(sizeof(unetk_head_bn_bwd_args) == 160, "ABI layout (see _lib.py)");

template <int DOUT>
struct HeadGrad {
  const float* dl;    // [N, DOUT, H, W]
  uint32_t hw;
  float w[DOUT][8];   // this thread's 8 channels of the head weight
  __device__ __forceinline__ void load_w(const float* wh, int c, int g) {
#pragma unroll
    for (int k = 0; k < DOUT; ++k) load8(wh + (size_t)k * c + g * 8, w[k]);
  }
  __device__ __forceinline__ void load_d(int64_t pix, float (&d)[DOUT]) const {
    const uint32_t p32 = (uint32_t)pix, img = p32 / hw, off = p32 - img * hw;
    const float* b = dl + ((size_t)img * DOUT) * hw + off;
#pragma unroll
    for (int k = 0; k < DOUT; ++k) d[k] = __ldg(b + (size_t)k * hw);
  }
};

// 16/32-byte raw loads kept in registers so that several pixels' loads are in flight before any is consumed
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { r = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

constexpr int kHeadUnroll = 4;   // pixels in flight per thread

// BatchNorm apply + ReLU without pooling, four 16-byte loads in flight per thread (the generic kernel above keeps one)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    bn_relu_apply_stream_kernel(const T* __restrict__ z, int zld, const float* __restrict__ scale,
                                const float* __restrict__ shift, T* __restrict__ a, int ald, int64_t npix, int cg) {
  const int64_t total = npix * cg;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;   // the host makes it a multiple of cg
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = (int)((uint32_t)first % (uint32_t)cg);
  const uint32_t pstride = (uint32_t)(stride / cg);
  float sc[8], sh[8];
  load8(scale + g * 8, sc);
  load8(shift + g * 8, sh);
  uint32_t p = (uint32_t)(first / cg);
  for (int64_t i0 = first; i0 < total; i0 += (int64_t)kHeadUnroll * stride, p += kHeadUnroll * pstride) {
    Raw8<T> raw[kHeadUnroll];
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u)
      if (i0 + (int64_t)u * stride < total) raw[u].load(z + (int64_t)(p + u * pstride) * zld + g * 8);
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      if (i0 + (int64_t)u * stride < total) {
        float v[8];
        raw[u].unpack(v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f);
        store8(a + (int64_t)(p + u * pstride) * ald + g * 8, v);
      }
    }
  }
}

// Reduction pass.  With mask = [z*scale+shift > 0] only two families of sums are accumulated per pixel,
//   M[k][c] = sum_p mask * dlogits[p][k]          Z[k][c] = sum_p mask * dlogits[p][k] * z[p][c]
// and everything else follows per channel at the end (a = mask * (z*scale + shift), da = sum_k dlogits[k] * Wh[k][c]):
//   sum dy   = sum_k Wh[k][c] * M[k][c]           sum dy*z = sum_k Wh[k][c] * Z[k][c]
//   dWh[k][c] = scale[c] * Z[k][c] + shift[c] * M[k][c]
template <typename T, int DOUT>
__global__ void __launch_bounds__(kThreads, 2)
    head_bn_bwd_reduce_kernel(const T* __restrict__ z, int zld, int c, int64_t npix, const float* __restrict__ scale,
                              const float* __restrict__ shift, const float* __restrict__ mean,
                              const float* __restrict__ invstd, HeadGrad<DOUT> hg, const float* __restrict__ wh, int cgb,
                              int items_per_block, double* __restrict__ sums) {
  const int rows = kThreads / cgb;
  const int lane_g = threadIdx.x % cgb, row = threadIdx.x / cgb;
  const int cg0 = blockIdx.x * cgb;
  const int g = cg0 + lane_g;
  const int64_t i0 = (int64_t)blockIdx.y * items_per_block, i1 = min(i0 + (int64_t)items_per_block, npix);
  float sc[8], sh[8];
  load8(scale + g * 8, sc);
  load8(shift + g * 8, sh);
  float M[DOUT][8], Z[DOUT][8], accb[DOUT];
#pragma unroll
  for (int k = 0; k < DOUT; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) M[k][j] = Z[k][j] = 0.f;
  }
  for (int64_t it = i0 + row; it < i1; it += (int64_t)kHeadUnroll * rows) {
    Raw8<T> raw[kHeadUnroll];
    float d[kHeadUnroll][DOUT];
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      const int64_t p = it + (int64_t)u * rows;
      if (p < i1) {
        raw[u].load(z + p * zld + g * 8);
        hg.load_d(p, d[u]);
      } else {
        raw[u].zero();
#pragma unroll
        for (int k = 0; k < DOUT; ++k) d[u][k] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      float zv[8];
      raw[u].unpack(zv);
#pragma unroll
      for (int k = 0; k < DOUT; ++k) accb[k] += d[u][k];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool pos = fmaf(zv[j], sc[j], sh[j]) > act_threshold<T>();
#pragma unroll
        for (int k = 0; k < DOUT; ++k) {
          const float m = pos ? d[u][k] : 0.f;
          M[k][j] += m;
          Z[k][j] = fmaf(m, zv[j], Z[k][j]);
        }
      }
    }
  }
  // rows of acc: sum dy | sum dy*xhat | dWh[k] (DOUT rows) | dbh (entries 0..DOUT-1 of the first channel group only)
  float acc[3 + DOUT][8];
  {
    float mu[8], is[8];
    load8(mean + g * 8, mu);
    load8(invstd + g * 8, is);
    hg.load_w(wh, c, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < DOUT; ++k) {
        s1 = fmaf(hg.w[k][j], M[k][j], s1);
        s2 = fmaf(hg.w[k][j], Z[k][j], s2);
        acc[2 + k][j] = fmaf(sc[j], Z[k][j], sh[j] * M[k][j]);
      }
      acc[0][j] = s1;
      acc[1][j] = is[j] * (s2 - mu[j] * s1);
      // bias gradient: the g == 0 column of threads sees every pixel of the block exactly once
      acc[2 + DOUT][j] = (g == 0 && j < DOUT) ? accb[j < DOUT ? j : 0] : 0.f;
    }
  }
  double* outs[3 + DOUT];
#pragma unroll
  for (int q = 0; q < 3 + DOUT; ++q) outs[q] = sums + (size_t)q * c;
  block_channel_reduce<3 + DOUT>(acc, cgb, row, lane_g, rows, cg0, outs);
}

template <typename T, int DOUT>
__global__ void __launch_bounds__(kThreads, 2)
    head_bn_bwd_apply_kernel(const T* __restrict__ z, int zld, int c, int64_t npix, const float* __restrict__ scale,
                             const float* __restrict__ shift, const float* __restrict__ mean,
                             const float* __restrict__ invstd, HeadGrad<DOUT> hg, const float* __restrict__ wh, int cgb,
                             int items_per_block, const double* __restrict__ sums, double inv_count, T* __restrict__ dz,
                             int dzld, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dwh,
                             float* __restrict__ dbh) {
  const int rows = kThreads / cgb;
  const int lane_g = threadIdx.x % cgb, row = threadIdx.x / cgb;
  const int g = blockIdx.x * cgb + lane_g;
  const int64_t i0 = (int64_t)blockIdx.y * items_per_block, i1 = min(i0 + (int64_t)items_per_block, npix);
  float sc[8], sh[8], cb[8], cc[8];
  load8(scale + g * 8, sc);
  load8(shift + g * 8, sh);
  hg.load_w(wh, c, g);
  {
    float mu[8], is[8];
    load8(mean + g * 8, mu);
    load8(invstd + g * 8, is);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m1 = (float)(sums[g * 8 + k] * inv_count);
      const float m2 = (float)(sums[c + g * 8 + k] * inv_count);
      cb[k] = -sc[k] * m2 * is[k];
      cc[k] = sc[k] * (m2 * is[k] * mu[k] - m1);
    }
  }
  if (blockIdx.y == 0 && row == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (dgamma) dgamma[g * 8 + k] = (float)sums[c + g * 8 + k];
      if (dbeta) dbeta[g * 8 + k] = (float)sums[g * 8 + k];
#pragma unroll
      for (int q = 0; q < DOUT; ++q) dwh[(size_t)q * c + g * 8 + k] = (float)sums[(size_t)(2 + q) * c + g * 8 + k];
    }
    if (g == 0 && dbh) {
#pragma unroll
      for (int q = 0; q < DOUT; ++q) dbh[q] = (float)sums[(size_t)(2 + DOUT) * c + q];
    }
  }
  // dz = scale*dy + cb*z + cc with dy = mask * sum_k dlogits[k]*Wh[k][c]: fold scale into the head weights
#pragma unroll
  for (int k = 0; k < DOUT; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) hg.w[k][j] *= sc[j];
  for (int64_t it = i0 + row; it < i1; it += (int64_t)kHeadUnroll * rows) {
    Raw8<T> raw[kHeadUnroll];
    float d[kHeadUnroll][DOUT];
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      const int64_t p = it + (int64_t)u * rows;
      if (p < i1) {
        raw[u].load(z + p * zld + g * 8);
        hg.load_d(p, d[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      const int64_t p = it + (int64_t)u * rows;
      if (p < i1) {
        float zv[8], o[8];
        raw[u].unpack(zv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float da = 0.f;
#pragma unroll
          for (int k = 0; k < DOUT; ++k) da = fmaf(d[u][k], hg.w[k][j], da);
          const float base = fmaf(cb[j], zv[j], cc[j]);
          o[j] = fmaf(zv[j], sc[j], sh[j]) > act_threshold<T>() ? base + da : base;
        }
        store8(dz + p * dzld + g * 8, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm apply + ReLU of the last block fused with the 1x1 classifier head (unet/unet.py:20-21 -> :91):
// logits[p][k] = bh[k] + sum_c Wh[k][c] * a[p][c],  a = T(relu(z*scale + shift)).  The C/8 threads that own one pixel
// are adjacent lanes, so the dot product finishes with log2(C/8) shuffles.  `a` itself is written only on request: with
// the fused head backward (head_bn_bwd_*) nothing reads it again, which saves a [N,H,W,C] write and the head's read.
// ------------------------------------------------------------------------------------------------
// C = 64 (8 channel groups): a warp owns 32 consecutive pixels.  Sub-iteration t covers pixels base + 4t + lane/8 with
// channel group lane%8 (512 contiguous bytes per load instruction, all eight loads issued before the first is used);
// the per-pixel dot products are finished with three shuffles and routed so that lane L ends up with the logits of
// pixel base + L, which makes the NCHW logits stores fully coalesced.
template <typename T, int DOUT>
__global__ void __launch_bounds__(kThreads)
    bn_relu_head_kernel(const T* __restrict__ z, int zld, const float* __restrict__ scale, const float* __restrict__ shift,
                        T* __restrict__ a, int ald, const float* __restrict__ wh, const float* __restrict__ bh,
                        float* __restrict__ logits, int64_t npix, uint32_t hw) {
  constexpr int C = 64;
  const int lane = threadIdx.x & 31, g = lane & 7, sub = lane >> 3;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float sc[8], sh[8], w[DOUT][8], b[DOUT];
  load8(scale + g * 8, sc);
  load8(shift + g * 8, sh);
#pragma unroll
  for (int k = 0; k < DOUT; ++k) {
    load8(wh + (size_t)k * C + g * 8, w[k]);
    b[k] = bh ? bh[k] : 0.f;
  }
  for (int64_t base = warp_id * 32; base < npix; base += nwarps * 32) {
    Raw8<T> raw[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int64_t p = base + 4 * t + sub;
      if (p < npix) raw[t].load(z + p * zld + g * 8); else raw[t].zero();
    }
    float out[DOUT];
#pragma unroll
    for (int k = 0; k < DOUT; ++k) out[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int64_t p = base + 4 * t + sub;
      float v[8];
      raw[t].unpack(v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = round_to<T>(fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f));
      if (a && p < npix) store8(a + p * ald + g * 8, v);
#pragma unroll
      for (int k = 0; k < DOUT; ++k) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(v[j], w[k][j], s);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        const float routed = __shfl_sync(0xffffffffu, s, (lane & 3) * 8);   // pixel base + 4t + (lane & 3)
        if ((lane >> 2) == t) out[k] = routed;
      }
    }
    const int64_t p = base + lane;
    if (p < npix) {
      const uint32_t p32 = (uint32_t)p, img = p32 / hw, off = p32 - img * hw;
#pragma unroll
      for (int k = 0; k < DOUT; ++k) logits[((size_t)img * DOUT + k) * hw + off] = out[k] + b[k];
    }
  }
}

// bf16 tier: the same fusion with the 64 x dout dot products on the tensor cores (mma.sync m16n8k16, fp32 accumulate).
// A warp owns 16 consecutive pixels per round.  Lane (gid = lane / 4, tig = lane % 4) loads, for pixels base + gid and
// base + gid + 8, the two 16-byte channel groups tig and tig + 4 (every load instruction reads 64 contiguous bytes per
// pixel row), applies BatchNorm + ReLU + bf16 rounding in registers and uses the packed words DIRECTLY as A fragments: the
// K slots of each of the four MMAs are a permutation of channels chosen so that a thread's own words are its fragment
// (slot 2*tig + {0,1} (+8) of MMA j  <->  channel 32*(j/2) + 8*tig + 2*(2*(j%2) + (slot >= 8)) + {0,1}); the B fragments
// (the head weights, constant) are built once per thread with the same permutation.  fp32 weights are split into
// hi + lo bf16 halves living in columns k and 4 + k of the N = 8 tile, so the product keeps fp32-level accuracy
// (|w - hi - lo| <= 2^-17 |w|); the two halves are added with one shuffle.  ~8 instructions per pixel instead of ~20.
__device__ __forceinline__ void mma_m16n8k16_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                  uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int DOUT>
__global__ void __launch_bounds__(kThreads, 3)
    bn_relu_head_mma_kernel(const __nv_bfloat16* __restrict__ z, int zld, const float* __restrict__ scale,
                            const float* __restrict__ shift, __nv_bfloat16* __restrict__ a, int ald,
                            const float* __restrict__ wh, const float* __restrict__ bh, float* __restrict__ logits,
                            int64_t npix, uint32_t hw) {
  constexpr int C = 64;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // this thread's 16 channels: groups tig (0..31) and tig + 4 (32..63)
  float sc[2][8], sh[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    load8(scale + (tig + 4 * h) * 8, sc[h]);
    load8(shift + (tig + 4 * h) * 8, sh[h]);
  }
  // B fragments: column n = gid (n < 4: hi half of Wh[n], n >= 4: lo half of Wh[n - 4]); MMA j, register r (K slots
  // 2*tig + {0,1} + 8*r) <-> channels 32*(j/2) + 8*tig + 2*(2*(j%2) + r) + {0,1}
  uint32_t bfrag[4][2];
  {
    const int k = gid & 3;
    const bool lo_half = gid >= 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int ch = 32 * (j >> 1) + 8 * tig + 2 * (2 * (j & 1) + r);
        float w0 = 0.f, w1 = 0.f;
        if (k < DOUT) {
          w0 = wh[(size_t)k * C + ch];
          w1 = wh[(size_t)k * C + ch + 1];
          if (lo_half) {
            w0 -= __bfloat162float(__float2bfloat16_rn(w0));
            w1 -= __bfloat162float(__float2bfloat16_rn(w1));
          }
        }
        bfrag[j][r] = pack_bf16x2(w0, w1);
      }
    }
  }
  // after the hi + lo fold, lane tig 0 holds classes 0,1 and lane tig 1 classes 2,3
  const float b0 = (bh && 2 * tig < DOUT) ? bh[2 * tig] : 0.f;
  const float b1 = (bh && 2 * tig + 1 < DOUT) ? bh[2 * tig + 1] : 0.f;
  BlockCursor cur(warp_id * 16, nwarps * 16, hw);
  for (int64_t base = warp_id * 16; base < npix; base += nwarps * 16, cur.advance()) {
    const int64_t p0 = base + gid, p1 = base + gid + 8;
    uint4 raw[2][2];                                 // [pixel row][channel half]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      raw[0][h] = p0 < npix ? *reinterpret_cast<const uint4*>(z + p0 * zld + (tig + 4 * h) * 8) : make_uint4(0, 0, 0, 0);
      raw[1][h] = p1 < npix ? *reinterpret_cast<const uint4*>(z + p1 * zld + (tig + 4 * h) * 8) : make_uint4(0, 0, 0, 0);
    }
    uint32_t act[2][2][4];                           // [pixel row][channel half][word = channel pair]
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t u[4] = {raw[r][h].x, raw[r][h].y, raw[r][h].z, raw[r][h].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lo = fmaxf(fmaf(__uint_as_float(u[i] << 16), sc[h][2 * i], sh[h][2 * i]), 0.f);
          const float hi = fmaxf(fmaf(__uint_as_float(u[i] & 0xffff0000u), sc[h][2 * i + 1], sh[h][2 * i + 1]), 0.f);
          act[r][h][i] = pack_bf16x2(lo, hi);
        }
      }
    }
    if (a) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (p0 < npix)
          *reinterpret_cast<uint4*>(a + p0 * ald + (tig + 4 * h) * 8) = make_uint4(act[0][h][0], act[0][h][1], act[0][h][2], act[0][h][3]);
        if (p1 < npix)
          *reinterpret_cast<uint4*>(a + p1 * ald + (tig + 4 * h) * 8) = make_uint4(act[1][h][0], act[1][h][1], act[1][h][2], act[1][h][3]);
      }
    }
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = j >> 1, w0 = 2 * (j & 1);
      mma_m16n8k16_bf16(c, act[0][h][w0], act[1][h][w0], act[0][h][w0 + 1], act[1][h][w0 + 1], bfrag[j][0], bfrag[j][1]);
    }
    // columns 4..7 (lanes tig 2,3) hold the lo-half products of classes 0..3: fold them onto lanes tig 0,1
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] += __shfl_down_sync(0xffffffffu, c[i], 2, 4);
    if (tig < 2) {
      const int k0 = 2 * tig;
      uint32_t img, off;
      if (p0 < npix) {
        cur.at(gid, img, off);
        if (k0 < DOUT) logits[((size_t)img * DOUT + k0) * hw + off] = c[0] + b0;
        if (k0 + 1 < DOUT) logits[((size_t)img * DOUT + k0 + 1) * hw + off] = c[1] + b1;
      }
      if (p1 < npix) {
        cur.at(gid + 8, img, off);
        if (k0 < DOUT) logits[((size_t)img * DOUT + k0) * hw + off] = c[2] + b0;
        if (k0 + 1 < DOUT) logits[((size_t)img * DOUT + k0 + 1) * hw + off] = c[3] + b1;
      }
    }
  }
}

// bf16 tier, C = 64: head backward apply with da = dlogits . (scale * Wh) on the tensor cores.  Same thread <-> data mapping
// idea as bn_relu_head_mma_kernel, transposed: here the MMA's M = 16 pixels, K = 16 holds the (class, hi/lo) terms and
// N = 8 a permuted set of channels, chosen so that the C fragments a thread receives are exactly the two 16-byte channel
// groups (tig and tig + 4) of the pixels (gid, gid + 8) it loads z for and stores dz to:
//   column n of N tile i  <->  channel 8 * G(n) + i,  G(2t) = t, G(2t + 1) = t + 4
//   K slots 0-3: hi(dl_k) * hi(W_k), 4-7: hi(dl_k) * lo(W_k), 8-11: lo(dl_k) * hi(W_k), 12-15: unused   (fp32-level accuracy)
template <int DOUT>
__global__ void __launch_bounds__(128, 3)
    head_bn_bwd_apply_mma_kernel(const __nv_bfloat16* __restrict__ z, int zld, int64_t npix, const float* __restrict__ scale,
                                 const float* __restrict__ shift, const float* __restrict__ mean,
                                 const float* __restrict__ invstd, const float* __restrict__ dl, uint32_t hw,
                                 const float* __restrict__ wh, const double* __restrict__ sums, double inv_count,
                                 __nv_bfloat16* __restrict__ dz, int dzld, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                 float* __restrict__ dwh, float* __restrict__ dbh) {
  constexpr int C = 64;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  if (blockIdx.x == 0 && threadIdx.x < C) {
    const int ch = threadIdx.x;
    if (dgamma) dgamma[ch] = (float)sums[C + ch];
    if (dbeta) dbeta[ch] = (float)sums[ch];
#pragma unroll
    for (int q = 0; q < DOUT; ++q) dwh[(size_t)q * C + ch] = (float)sums[(size_t)(2 + q) * C + ch];
    if (ch < DOUT && dbh) dbh[ch] = (float)sums[(size_t)(2 + DOUT) * C + ch];
  }
  // per-channel constants of this thread's 16 channels (groups tig and tig + 4)
  float sc[2][8], sh[2][8], cb[2][8], cc[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int g = tig + 4 * h;
    float mu[8], is[8];
    load8(scale + g * 8, sc[h]);
    load8(shift + g * 8, sh[h]);
    load8(mean + g * 8, mu);
    load8(invstd + g * 8, is);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m1 = (float)(sums[g * 8 + k] * inv_count);
      const float m2 = (float)(sums[C + g * 8 + k] * inv_count);
      cb[h][k] = -sc[h][k] * m2 * is[k];
      cc[h][k] = sc[h][k] * (m2 * is[k] * mu[k] - m1);
    }
  }
  // B fragments (constant): N tile i, column n = gid <-> channel 8 * G(gid) + i; register r covers K slots 2*tig + {0,1} + 8*r
  uint32_t bfrag[8][2];
  {
    const int grp = (gid >> 1) + 4 * (gid & 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = 8 * grp + i;
      const float s_ch = scale[ch];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int slot = 2 * tig + e + 8 * r, k = slot & 3, part = slot >> 2;   // part 0: hi(W), 1: lo(W), 2: hi(W), 3: unused
          float w = (k < DOUT && part < 3) ? wh[(size_t)k * C + ch] * s_ch : 0.f;
          if (part == 1) w -= __bfloat162float(__float2bfloat16_rn(w));
          v[e] = w;
        }
        bfrag[i][r] = pack_bf16x2(v[0], v[1]);
      }
    }
  }
  const int k0 = 2 * (tig & 1);                        // the two classes this thread feeds into the A fragments
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  BlockCursor cur(warp_id * 16, nwarps * 16, hw);
  for (int64_t base = warp_id * 16; base < npix; base += nwarps * 16, cur.advance()) {
    const int64_t p[2] = {base + gid, base + gid + 8};
    uint4 raw[2][2];
    float d[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const bool ok = p[r] < npix;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        raw[r][h] = ok ? *reinterpret_cast<const uint4*>(z + p[r] * zld + (tig + 4 * h) * 8) : make_uint4(0, 0, 0, 0);
      uint32_t img, off;
      cur.at(gid + 8 * r, img, off);
      const float* b = dl + ((size_t)img * DOUT) * hw + off;
      d[r][0] = (ok && k0 < DOUT) ? __ldg(b + (size_t)k0 * hw) : 0.f;
      d[r][1] = (ok && k0 + 1 < DOUT) ? __ldg(b + (size_t)(k0 + 1) * hw) : 0.f;
    }
    // A fragments: rows = pixels, K slots 2*tig + {0,1}: hi(dl) (slots 0-7); slots 8 + 2*tig + {0,1}: lo(dl) for tig < 2
    uint32_t afr[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float h0 = __bfloat162float(__float2bfloat16_rn(d[r][0])), h1 = __bfloat162float(__float2bfloat16_rn(d[r][1]));
      afr[r] = pack_bf16x2(h0, h1);
      afr[2 + r] = tig < 2 ? pack_bf16x2(d[r][0] - h0, d[r][1] - h1) : 0u;
    }
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      mma_m16n8k16_bf16(acc[i], afr[0], afr[1], afr[2], afr[3], bfrag[i][0], bfrag[i][1]);
    }
    // acc[i][2*r + h] = da of pixel p[r], channel 8 * (tig + 4*h) + i
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (p[r] < npix) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t u[4] = {raw[r][h].x, raw[r][h].y, raw[r][h].z, raw[r][h].w};
          uint32_t o[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float z0 = __uint_as_float(u[w] << 16), z1 = __uint_as_float(u[w] & 0xffff0000u);
            const int i0 = 2 * w, i1 = 2 * w + 1;
            float v0 = fmaf(cb[h][i0], z0, cc[h][i0]), v1 = fmaf(cb[h][i1], z1, cc[h][i1]);
            if (fmaf(z0, sc[h][i0], sh[h][i0]) > act_threshold<__nv_bfloat16>()) v0 += acc[i0][2 * r + h];
            if (fmaf(z1, sc[h][i1], sh[h][i1]) > act_threshold<__nv_bfloat16>()) v1 += acc[i1][2 * r + h];
            o[w] = pack_bf16x2(v0, v1);
          }
          *reinterpret_cast<uint4*>(dz + p[r] * dzld + (tig + 4 * h) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// bf16 tier, C = 64: the reduction pass with the per-class sums on the tensor cores.  M[k][c] = sum_p mask * dl[p][k] and
// Z[k][c] = sum_p mask * dl[p][k] * z[p][c] are matrix products over the PIXEL dimension (K = the 16 pixels of a warp's
// block): A = the mask as 1.0 / 0 resp. mask * z (both exact in bf16) with 16 channels as rows, B = dl with the 8 columns
// (class k, hi / lo half of the fp32 value: fp32-level accuracy).  Lane (gid, tig) reads the 16-byte channel group `gid` of
// the four pixels 2*tig + {0,1,8,9} -- exactly the K slots of its fragments -- and M tile t takes channels 8*gid + 2t (row
// gid) and 8*gid + 2t + 1 (row gid + 8), so two byte-permutes per register build the A fragments from the masked words.
// The 32 accumulators live in registers for the whole grid-stride loop; hi + lo columns meet through one shuffle, warps
// through shared memory in a fixed order, blocks through fp64 atomics.  z and dl reach the warp through a private
// 4-stage cp.async ring (a 16-pixel block = 2 KB of z + 16 * dout floats), so three blocks per warp are in flight while one
// is being reduced: the loop is issue-bound, not latency-bound.
struct HeadStage {
  uint4 z[16][8];                                      // [pixel][channel group ^ swizzle]
  float d[4][16];                                      // [class][pixel]
};
constexpr int kHeadStages = 4;

// Producer side of the ring: the warp copies one 16-pixel block per round.  Lane l moves the 16-byte chunks l + 32 r
// (pixel (l >> 3) + 4 r, channel group l & 7) and the dl values of pixel l & 15 for the classes (l >> 4) and (l >> 4) + 2.
// Whole blocks that do not straddle two images (hw % 16 == 0: every real head) take the short path: a running source
// pointer and per-lane destination offsets fixed before the loop; the tail block and odd plane sizes take the general one.
template <int DOUT>
struct HeadFill {
  const char* zsrc;                                    // this lane's first chunk of the block to fetch next
  int64_t zstep, base, stride, npix;
  uint32_t zrow4, zdst[4], ddst;                       // bytes between the chunks of a lane; offsets inside a stage
  const float* dl;
  const __nv_bfloat16* z;
  int zld, lane;
  bool aligned;
  BlockCursor cur;
  __device__ __forceinline__ HeadFill(const __nv_bfloat16* z_, int zld_, const float* dl_, int64_t base_, int64_t stride_,
                                      int64_t npix_, uint32_t hw, int lane_)
      : base(base_), stride(stride_), npix(npix_), dl(dl_), z(z_), zld(zld_), lane(lane_), cur(base_, stride_, hw) {
    zsrc = reinterpret_cast<const char*>(z_ + (base_ + (lane_ >> 3)) * zld_ + (lane_ & 7) * 8);
    zstep = stride_ * zld_ * 2;
    zrow4 = 4u * (uint32_t)zld_ * 2u;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int px = (lane_ >> 3) + 4 * r, g = lane_ & 7;
      zdst[r] = (uint32_t)(px * 128 + ((g ^ (2 * ((px >> 1) & 3))) * 16));
    }
    ddst = (uint32_t)(sizeof(uint4) * 16 * 8 + ((lane_ >> 4) * 16 + (lane_ & 15)) * 4);
    aligned = (hw & 15u) == 0;
  }
  __device__ __forceinline__ void issue(uint32_t stage_addr) {
    if (base < npix) {
      if (aligned && base + 16 <= npix) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(stage_addr + zdst[r]), "l"(zsrc + (size_t)r * zrow4));
        const float* src = dl + ((size_t)cur.img * DOUT + (lane >> 4)) * cur.hw + cur.off + (lane & 15);
        if ((lane >> 4) < DOUT)                          // dout = 1: the upper half-warp has no class to fetch
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(stage_addr + ddst), "l"(src));
        if ((lane >> 4) + 2 < DOUT)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(stage_addr + ddst + 128u), "l"(src + 2 * (size_t)cur.hw));
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int64_t p = base + (lane >> 3) + 4 * r;
          const bool ok = p < npix;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(stage_addr + zdst[r]),
                       "l"(z + (ok ? p : 0) * zld + (lane & 7) * 8), "r"(ok ? 16 : 0));
        }
        const int t = lane & 15;
        const bool ok = base + t < npix;
        uint32_t img = 0, off = 0;
        if (ok) cur.at(t, img, off);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int k = (lane >> 4) + 2 * r;
          if (k < DOUT)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(stage_addr + ddst + 128u * r),
                         "l"(dl + ((size_t)img * DOUT + k) * cur.hw + off), "r"(ok ? 4 : 0));
        }
      }
    }
    cp_async_commit();
    base += stride;
    zsrc += zstep;
    cur.advance();
  }
};

template <int DOUT>
__global__ void __launch_bounds__(128, 4)
    head_bn_bwd_reduce_mma_kernel(const __nv_bfloat16* __restrict__ z, int zld, int64_t npix, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ mean,
                                  const float* __restrict__ invstd, const float* __restrict__ dl, uint32_t hw,
                                  const float* __restrict__ wh, double* __restrict__ sums) {
  constexpr int C = 64, S = kHeadStages;
  __shared__ __align__(16) HeadStage ring_mem[4 * S];
  static_assert(sizeof(HeadStage) * 4 * S >= sizeof(float) * 4 * (2 * 4 + 1) * C, "the warp partials reuse the ring");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
  HeadStage* ring = ring_mem + warp * S;
  const int k = gid & 3;
  const bool lo_col = gid >= 4;
  float sc[8], sh[8];
  load8(scale + gid * 8, sc);
  load8(shift + gid * 8, sh);
  float accM[4][4], accZ[4][4], accb = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int q = 0; q < 4; ++q) accM[t][q] = accZ[t][q] = 0.f;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t stride = nwarps * 16;
  // producer side of the ring runs S - 1 blocks ahead of the consumer
  HeadFill<DOUT> fill(z, zld, dl, warp_id * 16, stride, npix, hw, lane);
  const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring);
#pragma unroll
  for (int s = 0; s < S - 1; ++s) fill.issue(ring_addr + s * (uint32_t)sizeof(HeadStage));
  int stage = 0;
  for (int64_t base = warp_id * 16; base < npix; base += stride) {
    // refill the stage the previous round finished reading
    fill.issue(ring_addr + (uint32_t)(stage == 0 ? S - 1 : stage - 1) * (uint32_t)sizeof(HeadStage));
    cp_async_wait<S - 1>();
    __syncwarp();
    const HeadStage& st = ring[stage];
    uint32_t zw[4][4], ow[4][4];                       // masked z words / mask-as-1.0 words, [pixel slot][channel pair]
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int px = 2 * tig + (j & 1) + 8 * (j >> 1);
      const uint4 raw = st.z[px][gid ^ (2 * tig)];
      d[j] = k < DOUT ? st.d[k][px] : 0.f;
      const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const bool m0 = fmaf(__uint_as_float(u[w] << 16), sc[2 * w], sh[2 * w]) > act_threshold<__nv_bfloat16>();
        const bool m1 = fmaf(__uint_as_float(u[w] & 0xffff0000u), sc[2 * w + 1], sh[2 * w + 1]) > act_threshold<__nv_bfloat16>();
        const uint32_t mb = (m0 ? 0x0000ffffu : 0u) | (m1 ? 0xffff0000u : 0u);
        zw[j][w] = u[w] & mb;
        ow[j][w] = 0x3f803f80u & mb;
      }
    }
    if (!lo_col) accb += (d[0] + d[1]) + (d[2] + d[3]);
    // B fragment: column gid = (class k, hi / lo half of dl); K slots = this thread's four pixels
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float h = __bfloat162float(__float2bfloat16_rn(d[j]));
      f[j] = lo_col ? d[j] - h : h;
    }
    const uint32_t b0 = pack_bf16x2(f[0], f[1]), b1 = pack_bf16x2(f[2], f[3]);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      mma_m16n8k16_bf16(accZ[t], __byte_perm(zw[0][t], zw[1][t], 0x5410u), __byte_perm(zw[0][t], zw[1][t], 0x7632u),
                        __byte_perm(zw[2][t], zw[3][t], 0x5410u), __byte_perm(zw[2][t], zw[3][t], 0x7632u), b0, b1);
      mma_m16n8k16_bf16(accM[t], __byte_perm(ow[0][t], ow[1][t], 0x5410u), __byte_perm(ow[0][t], ow[1][t], 0x7632u),
                        __byte_perm(ow[2][t], ow[3][t], 0x5410u), __byte_perm(ow[2][t], ow[3][t], 0x7632u), b0, b1);
    }
    __syncwarp();                                      // every lane is done with this stage before it is refilled
    stage = stage + 1 == S ? 0 : stage + 1;
  }
  cp_async_wait<0>();
  // columns 2*tig + {0,1}: tig 0,1 hold the hi halves of classes (0,1) / (2,3), tig 2,3 the lo halves: fold
#pragma unroll
  for (int t = 0; t < 4; ++t) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      accM[t][q] += __shfl_down_sync(0xffffffffu, accM[t][q], 2);
      accZ[t][q] += __shfl_down_sync(0xffffffffu, accZ[t][q], 2);
    }
  }
  // bias gradient: sum over the four tig lanes of a class column (only the hi columns counted their pixels)
  accb += __shfl_xor_sync(0xffffffffu, accb, 1);
  accb += __shfl_xor_sync(0xffffffffu, accb, 2);
  __syncthreads();                                     // all warps are done with their rings: reuse them for the partials
  float(*part)[2 * DOUT + 1][C] = reinterpret_cast<float(*)[2 * DOUT + 1][C]>(ring_mem);   // [warp][M_k | Z_k | db][channel]
  if (tig < 2) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ch = 8 * gid + 2 * t + (q >> 1), kk = 2 * tig + (q & 1);
        if (kk < DOUT) {
          part[warp][kk][ch] = accM[t][q];
          part[warp][DOUT + kk][ch] = accZ[t][q];
        }
      }
    }
  }
  if (gid < DOUT && tig == 0) part[warp][2 * DOUT][gid] = accb;
  __syncthreads();
  if (threadIdx.x < C) {
    const int ch = threadIdx.x;
    float M[DOUT], Z[DOUT];
#pragma unroll
    for (int q = 0; q < DOUT; ++q) {
      M[q] = (part[0][q][ch] + part[1][q][ch]) + (part[2][q][ch] + part[3][q][ch]);
      Z[q] = (part[0][DOUT + q][ch] + part[1][DOUT + q][ch]) + (part[2][DOUT + q][ch] + part[3][DOUT + q][ch]);
    }
    const float scc = scale[ch], shc = shift[ch], mu = mean[ch], is = invstd[ch];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < DOUT; ++q) {
      const float w = wh[(size_t)q * C + ch];
      s1 = fmaf(w, M[q], s1);
      s2 = fmaf(w, Z[q], s2);
      atomicAdd(sums + (size_t)(2 + q) * C + ch, (double)fmaf(scc, Z[q], shc * M[q]));
    }
    atomicAdd(sums + ch, (double)s1);
    atomicAdd(sums + C + ch, (double)(is * (s2 - mu * s1)));
    if (ch < DOUT) {
      const float db = (part[0][2 * DOUT][ch] + part[1][2 * DOUT][ch]) + (part[2][2 * DOUT][ch] + part[3][2 * DOUT][ch]);
      atomicAdd(sums + (size_t)(2 + DOUT) * C + ch, (double)db);
    }
  }
}

static int check_head_bn(const unetk_head_bn_bwd_args* a) {
  UNETK_REQUIRE(a != nullptr, "head_bn_bwd: null args");
  UNETK_REQUIRE(tensor_ok(a->z) && vec8_ok(a->z), "head_bn_bwd: z must be NHWC with c%%8==0, ld%%8==0, 16B aligned");
  UNETK_REQUIRE(a->dlogits && a->w_head && a->scale && a->shift && a->mean && a->invstd && a->sums,
                "head_bn_bwd: null pointer");
  UNETK_REQUIRE(pixels(a->z) < (1LL << 31), "head_bn_bwd: more than 2^31 pixels");
  if (a->dout < 1 || a->dout > 4) {
    set_error("head_bn_bwd: fused path supports 1..4 classes, got %d (use unetk_head_bwd + unetk_bn_relu_bwd_*)", a->dout);
    return UNETK_ERR_UNSUPPORTED;
  }
  return UNETK_OK;
}

#define UNETK_DISPATCH_DOUT(dout, D, ...)   \
  switch (dout) {                           \
    case 1: { constexpr int D = 1; __VA_ARGS__ } break; \
    case 2: { constexpr int D = 2; __VA_ARGS__ } break; \
    case 3: { constexpr int D = 3; __VA_ARGS__ } break; \
    default: { constexpr int D = 4; __VA_ARGS__ } break; \
  }

}  // namespace unetk

using namespace unetk;

extern "C" {

int unetk_bn_stats(const unetk_tensor* z, double* sum, double* sumsq, void* stream) {
  UNETK_REQUIRE(z && sum && sumsq, "bn_stats: null argument");
  UNETK_REQUIRE(tensor_ok(*z) && vec8_ok(*z), "bn_stats: z must be NHWC with c%%8==0, ld%%8==0, 16B aligned");
  const int cg = z->c / 8, cgb = choose_cgb(cg), rows = kThreads / cgb;
  const int64_t npix = pixels(*z);
  int ppb = rows * 32;
  const int64_t nb = (npix + ppb - 1) / ppb;
  UNETK_REQUIRE(nb <= 65535 * 16, "bn_stats: tensor too large");
  if (nb > 65535) ppb *= 16;
  dim3 grid(cg / cgb, (unsigned)((npix + ppb - 1) / ppb));
  const size_t smem = (size_t)2 * rows * cgb * 8 * sizeof(float);
  UNETK_DISPATCH_DTYPE(z->dtype, T, {
    bn_stats_kernel<T><<<grid, kThreads, smem, (cudaStream_t)stream>>>((const T*)z->ptr, npix, z->ld, cgb, ppb, sum, sumsq);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_bn_finalize(const unetk_bn_finalize_args* a, void* stream) {
  UNETK_REQUIRE(a && a->c > 0 && a->gamma && a->beta && a->scale && a->shift, "bn_finalize: null argument");
  if (a->training) {
    UNETK_REQUIRE(a->sum && a->sumsq && a->count > 0 && a->mean && a->invstd, "bn_finalize: training needs sum/sumsq/count/mean/invstd");
  } else {
    UNETK_REQUIRE(a->running_mean && a->running_var, "bn_finalize: eval needs running statistics");
  }
  bn_finalize_kernel<<<(a->c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*a);
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_bn_relu_apply(const unetk_tensor* z, const float* scale, const float* shift, const unetk_tensor* a,
                        const unetk_tensor* pooled, uint16_t* pool_idx, void* stream) {
  UNETK_REQUIRE(z && a && scale && shift, "bn_relu_apply: null argument");
  UNETK_REQUIRE(tensor_ok(*z) && vec8_ok(*z) && tensor_ok(*a) && vec8_ok(*a), "bn_relu_apply: bad z/a tensor");
  UNETK_REQUIRE(a->dtype == z->dtype && a->n == z->n && a->h == z->h && a->w == z->w && a->c == z->c,
                "bn_relu_apply: a must match z");
  const bool pool = pooled && pooled->ptr;
  if (pool) {
    UNETK_REQUIRE(tensor_ok(*pooled) && vec8_ok(*pooled) && pooled->dtype == z->dtype && pooled->n == z->n &&
                      pooled->h * 2 == z->h && pooled->w * 2 == z->w && pooled->c == z->c,
                  "bn_relu_apply: pooled must be [N,H/2,W/2,C]");
  }
  const int cg = z->c / 8;
  const int64_t items = pixels(*z) / (pool ? 4 : 1) * cg;
  UNETK_REQUIRE(items < (1LL << 32), "bn_relu_apply: tensor too large");
  const int grid = grid_for(items);
  UNETK_DISPATCH_DTYPE(z->dtype, T, {
    if (pool)
      bn_relu_apply_kernel<T, true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)z->ptr, z->ld, scale, shift, (T*)a->ptr, a->ld, (T*)pooled->ptr, pooled->ld, pool_idx, z->n, z->h, z->w, cg);
    else if ((kThreads % cg) == 0 && pixels(*z) < (1LL << 31))
      bn_relu_apply_stream_kernel<T><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)z->ptr, z->ld, scale, shift, (T*)a->ptr, a->ld, pixels(*z), cg);
    else
      bn_relu_apply_kernel<T, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)z->ptr, z->ld, scale, shift, (T*)a->ptr, a->ld, nullptr, 0, nullptr, z->n, z->h, z->w, cg);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_bn_relu_bwd_reduce(const unetk_bn_bwd_args* a, void* stream) {
  int rc = check_bwd(a);
  if (rc) return rc;
  const bool pool = a->dpool.ptr != nullptr;
  const int cg = a->z.c / 8, cgb = choose_cgb(cg), rows = kThreads / cgb;
  const int64_t nitems = pixels(a->z) / (pool ? 2 : 1);
  int ipb = rows * (pool ? 16 : 32);
  if ((nitems + ipb - 1) / ipb > 65535) ipb *= 16;
  UNETK_REQUIRE((nitems + ipb - 1) / ipb <= 65535, "bn_bwd_reduce: tensor too large");
  dim3 grid(cg / cgb, (unsigned)((nitems + ipb - 1) / ipb));
  const size_t smem = (size_t)2 * rows * cgb * 8 * sizeof(float);
  double* s1 = a->sums;
  double* s2 = a->sums + a->z.c;
  UNETK_DISPATCH_DTYPE(a->z.dtype, T, {
    BwdSrc<T> s = make_src<T>(a);
    if (pool)
      bn_bwd_reduce_kernel<T, true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(s, cgb, ipb, s1, s2);
    else
      bn_bwd_reduce_kernel<T, false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(s, cgb, ipb, s1, s2);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_bn_relu_bwd_apply(const unetk_bn_bwd_args* a, void* stream) {
  int rc = check_bwd(a);
  if (rc) return rc;
  UNETK_REQUIRE(tensor_ok(a->dz) && vec8_ok(a->dz) && a->dz.dtype == a->z.dtype && a->dz.n == a->z.n &&
                    a->dz.h == a->z.h && a->dz.w == a->z.w && a->dz.c == a->z.c, "bn_bwd_apply: dz must match z");
  const bool pool = a->dpool.ptr != nullptr;
  UNETK_REQUIRE(pool || a->dy.ptr, "bn_bwd_apply: dy required without dpool");
  const int cg = a->z.c / 8, cgb = choose_cgb(cg), rows = kThreads / cgb;
  const int64_t nitems = pixels(a->z) / (pool ? 2 : 1);
  int ipb = rows * (pool ? 8 : 16);
  if ((nitems + ipb - 1) / ipb > 65535) ipb *= 16;
  UNETK_REQUIRE((nitems + ipb - 1) / ipb <= 65535, "bn_bwd_apply: tensor too large");
  dim3 grid(cg / cgb, (unsigned)((nitems + ipb - 1) / ipb));
  const double inv_count = 1.0 / (double)pixels(a->z);
  const double* s1 = a->sums;
  const double* s2 = a->sums + a->z.c;
  UNETK_DISPATCH_DTYPE(a->z.dtype, T, {
    BwdSrc<T> s = make_src<T>(a);
    if (pool)
      bn_bwd_apply_kernel<T, true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(s, cgb, ipb, s1, s2, inv_count, (T*)a->dz.ptr, a->dz.ld, a->dgamma, a->dbeta);
    else
      bn_bwd_apply_kernel<T, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(s, cgb, ipb, s1, s2, inv_count, (T*)a->dz.ptr, a->dz.ld, a->dgamma, a->dbeta);
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_head_bn_bwd_reduce(const unetk_head_bn_bwd_args* a, void* stream) {
  int rc = check_head_bn(a);
  if (rc) return rc;
  const int cg = a->z.c / 8, cgb = choose_cgb(cg), rows = kThreads / cgb;
  const int64_t npix = pixels(a->z);
  int ipb = rows * 64;
  if ((npix + ipb - 1) / ipb > 65535) ipb *= 16;
  UNETK_REQUIRE((npix + ipb - 1) / ipb <= 65535, "head_bn_bwd_reduce: tensor too large");
  dim3 grid(cg / cgb, (unsigned)((npix + ipb - 1) / ipb));
  if (a->z.dtype == UNETK_BF16 && a->z.c == 64) {
    // tensor-core formulation (the U-Net head: 64 channels): persistent grid, 16 pixels per warp and round
    const int64_t warps = (npix + 15) / 16;
    int64_t blocks = (warps + 3) / 4;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    UNETK_DISPATCH_DOUT(a->dout, D, {
      head_bn_bwd_reduce_mma_kernel<D><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)a->z.ptr, a->z.ld, npix, a->scale, a->shift, a->mean, a->invstd, a->dlogits,
          (uint32_t)(a->z.h * a->z.w), a->w_head, a->sums);
    });
    UNETK_LAUNCH_CHECK();
    return UNETK_OK;
  }
  UNETK_DISPATCH_DTYPE(a->z.dtype, T, {
    UNETK_DISPATCH_DOUT(a->dout, D, {
      HeadGrad<D> hg;
      hg.dl = a->dlogits;
      hg.hw = (uint32_t)(a->z.h * a->z.w);
      const size_t smem = (size_t)(3 + D) * rows * cgb * 8 * sizeof(float);
      if (smem > 48 * 1024)
        UNETK_CUDA(cudaFuncSetAttribute(head_bn_bwd_reduce_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      head_bn_bwd_reduce_kernel<T, D><<<grid, kThreads, smem, (cudaStream_t)stream>>>(
          (const T*)a->z.ptr, a->z.ld, a->z.c, npix, a->scale, a->shift, a->mean, a->invstd, hg, a->w_head, cgb, ipb, a->sums);
    });
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_head_bn_bwd_apply(const unetk_head_bn_bwd_args* a, void* stream) {
  int rc = check_head_bn(a);
  if (rc) return rc;
  UNETK_REQUIRE(tensor_ok(a->dz) && vec8_ok(a->dz) && a->dz.dtype == a->z.dtype && a->dz.n == a->z.n &&
                    a->dz.h == a->z.h && a->dz.w == a->z.w && a->dz.c == a->z.c, "head_bn_bwd_apply: dz must match z");
  UNETK_REQUIRE(a->dw_head != nullptr, "head_bn_bwd_apply: dw_head required");
  const int cg = a->z.c / 8, cgb = choose_cgb(cg), rows = kThreads / cgb;
  const int64_t npix = pixels(a->z);
  int ipb = rows * 16;
  if ((npix + ipb - 1) / ipb > 65535) ipb *= 16;
  UNETK_REQUIRE((npix + ipb - 1) / ipb <= 65535, "head_bn_bwd_apply: tensor too large");
  dim3 grid(cg / cgb, (unsigned)((npix + ipb - 1) / ipb));
  const double inv_count = 1.0 / (double)npix;
  if (a->z.dtype == UNETK_BF16 && a->z.c == 64) {
    // tensor-core formulation (the U-Net head: 64 channels): 16 pixels per warp and round
    const int64_t warps = (npix + 15) / 16;
    int64_t blocks = (warps * 32 + 127) / 128;
    const int64_t cap = (int64_t)sm_count() * 12;
    if (blocks > cap) blocks = cap;
    UNETK_DISPATCH_DOUT(a->dout, D, {
      head_bn_bwd_apply_mma_kernel<D><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)a->z.ptr, a->z.ld, npix, a->scale, a->shift, a->mean, a->invstd, a->dlogits,
          (uint32_t)(a->z.h * a->z.w), a->w_head, a->sums, inv_count, (__nv_bfloat16*)a->dz.ptr, a->dz.ld, a->dgamma, a->dbeta,
          a->dw_head, a->db_head);
    });
    UNETK_LAUNCH_CHECK();
    return UNETK_OK;
  }
  UNETK_DISPATCH_DTYPE(a->z.dtype, T, {
    UNETK_DISPATCH_DOUT(a->dout, D, {
      HeadGrad<D> hg;
      hg.dl = a->dlogits;
      hg.hw = (uint32_t)(a->z.h * a->z.w);
      head_bn_bwd_apply_kernel<T, D><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)a->z.ptr, a->z.ld, a->z.c, npix, a->scale, a->shift, a->mean, a->invstd, hg, a->w_head, cgb, ipb, a->sums,
          inv_count, (T*)a->dz.ptr, a->dz.ld, a->dgamma, a->dbeta, a->dw_head, a->db_head);
    });
  });
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}

int unetk_bn_relu_head_fprop(const unetk_tensor* z, const float* scale, const float* shift, const unetk_tensor* a,
                             const float* w_head, const float* b_head, int32_t dout, float* logits_nchw, void* stream) {
  UNETK_REQUIRE(z && scale && shift && w_head && logits_nchw, "bn_relu_head_fprop: null argument");
  UNETK_REQUIRE(tensor_ok(*z) && vec8_ok(*z), "bn_relu_head_fprop: bad z tensor");
  const bool store_a = a && a->ptr;
  if (store_a)
    UNETK_REQUIRE(tensor_ok(*a) && vec8_ok(*a) && a->dtype == z->dtype && a->n == z->n && a->h == z->h && a->w == z->w &&
                      a->c == z->c, "bn_relu_head_fprop: a must match z");
  if (dout < 1 || dout > 4 || z->c != 64) {
    set_error("bn_relu_head_fprop: fused path needs 1..4 classes and C == 64 (use unetk_bn_relu_apply + unetk_head_fprop)");
    return UNETK_ERR_UNSUPPORTED;
  }
  UNETK_REQUIRE(pixels(*z) < (1LL << 31), "bn_relu_head_fprop: more than 2^31 pixels");
  const int64_t npix = pixels(*z);
  const int grid = grid_for(npix);   // one thread per pixel: a warp owns 32 pixels per round
  if (z->dtype == UNETK_BF16) {
    // tensor-core formulation of the dot products (16 pixels per warp and round)
    const int grid_mma = grid_for((npix + 15) / 16 * 32);
    UNETK_DISPATCH_DOUT(dout, D, {
      bn_relu_head_mma_kernel<D><<<grid_mma, kThreads, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)z->ptr, z->ld, scale, shift, store_a ? (__nv_bfloat16*)a->ptr : nullptr, store_a ? a->ld : 0,
          w_head, b_head, logits_nchw, npix, (uint32_t)(z->h * z->w));
    });
  } else {
    UNETK_DISPATCH_DOUT(dout, D, {
      bn_relu_head_kernel<float, D><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
          (const float*)z->ptr, z->ld, scale, shift, store_a ? (float*)a->ptr : nullptr, store_a ? a->ld : 0, w_head, b_head,
          logits_nchw, npix, (uint32_t)(z->h * z->w));
    });
  }
  UNETK_LAUNCH_CHECK();
  return UNETK_OK;
}
}
