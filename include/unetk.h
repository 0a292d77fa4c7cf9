/*
 * unetk.h -- C ABI of the B200-native U-Net training-step kernels (libunetk.so).
 *
 * This is the drop-in boundary of the hot path.  The reference (in5omnia/Image_Segmentation)
 * has no FFI of its own: its hot path is a chain of PyTorch ATen calls made from
 *   unet/unet.py:16-21,40,59,63,91      (conv3x3, BatchNorm, ReLU, MaxPool, ConvTranspose, cat, 1x1 head)
 *   utils/weighted_loss.py:36-98,163    (softmax, one-hot, Dice sums, cross entropy)
 *   utils/MetricsHistory.py:65-86       (argmax, one-hot, TP/FP/FN/TN)
 * Each entry point below names the reference call site(s) it replaces.  The host side
 * (image_segmentation_b200/, Python, mirrors the reference's nn.Module / loss / metrics API)
 * binds these with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs.  No torch / C++ types.
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator); the library
 *     never allocates, frees or keeps device memory beyond the call.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised.
 *   - return 0 on success, negative on error; unetk_last_error() gives the thread-local message.
 *   - activations are NHWC ("pixel-major"): element (n,h,w,c) lives at ptr[((n*H+h)*W+w)*ld + c];
 *     `ld` >= c lets a tensor be a channel slice of a wider buffer (zero-copy skip concatenation).
 *   - dtype: UNETK_F32 (parity tier, SIMT kernels) or UNETK_BF16 (throughput tier, tcgen05 kernels,
 *     fp32 accumulation).  There is no CPU fallback anywhere.
 */
#ifndef UNETK_H
#define UNETK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The shared library is built with -fvisibility=hidden: only the entry points declared here are exported. */
#if defined(__GNUC__)
#define UNETK_API __attribute__((visibility("default")))
#else
#define UNETK_API
#endif

#define UNETK_F32 0
#define UNETK_BF16 1
#define UNETK_U8 2   /* label buffers only */
#define UNETK_I64 3  /* label buffers only */

#define UNETK_OK 0
#define UNETK_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define UNETK_ERR_CUDA (-2)      /* CUDA runtime / driver error */
#define UNETK_ERR_UNSUPPORTED (-3)

/* algo selector for the contraction kernels */
#define UNETK_ALGO_AUTO 0   /* bf16 with channel counts multiple of 64 -> tcgen05; everything else -> SIMT */
#define UNETK_ALGO_SIMT 1   /* CUDA-core FFMA kernels (any dtype; the fp32 parity tier) */
#define UNETK_ALGO_TC 2     /* TMA + tcgen05.mma + TMEM (bf16 only) */
#define UNETK_ALGO_MASK 0xff
/* Experiment switches for the tcgen05 path, OR-ed into `algo` (the library reads no environment variables):
 * each one turns OFF a default optimisation so that its effect can be measured (tools/run_layer.py). */
#define UNETK_TC_NO_PAIR (1 << 8)        /* single-CTA kernels instead of cta_group::2 pairs */
#define UNETK_TC_NO_HALO (1 << 9)        /* one TMA box per tap instead of one halo tile per 64-channel chunk */
#define UNETK_TC_NO_EVEN_GROUPS (1 << 10) /* do not trim the pair count to a multiple of the N-tile count */
#define UNETK_TC_NO_HALO_N256 (1 << 11)  /* halo reuse only for N tiles <= 128 */
#define UNETK_TC_NO_HALO_PAIR (1 << 12)  /* halo kernel without CTA pairs */
#define UNETK_TC_NO_WGRAD_C64 (1 << 13)  /* generic 3-tap weight-gradient kernel for 64-channel gradients */
/* unetk_wgrad only: run-to-run reproducible weight gradients.  The pixel reduction is split over CTAs; by default the
 * splits meet in dw through fp32 atomics (order varies run to run).  With this flag every split STORES its partial result
 * to the caller's `partial` buffer and a second kernel adds the splits in index order. */
#define UNETK_TC_DETERMINISTIC (1 << 14)

typedef struct unetk_tensor {
  void* ptr;
  int32_t n, h, w, c;
  int32_t ld;    /* elements between consecutive pixels (>= c) */
  int32_t dtype; /* UNETK_F32 | UNETK_BF16 */
} unetk_tensor;

/* ---- library ---------------------------------------------------------------------------- */
UNETK_API int unetk_version(void);
UNETK_API const char* unetk_last_error(void);
/* sm_count / compute capability of the current device */
UNETK_API int unetk_device_query(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* Caller-provided scratch sizes.  The library never allocates device memory; every accumulator / workspace the
 * entry points below take is allocated (and, where stated, zeroed) by the caller.  unetk_query_workspace returns
 * the size in BYTES of buffer `what` for a problem described by up to four integers, or a negative error code.
 *   UNETK_WS_BN_STATS     (C)            stat_sum + stat_sumsq of unetk_conv / unetk_bn_stats        2*C doubles
 *   UNETK_WS_BN_BWD_SUMS  (C)            unetk_bn_bwd_args.sums / unetk_conv_args.bn_sums            2*C doubles
 *   UNETK_WS_HEAD_BN_SUMS (C, dout)      unetk_head_bn_bwd_args.sums                                 (3+dout)*C doubles
 *   UNETK_WS_POOL_IDX     (N, H, W, C)   pool_idx of unetk_bn_relu_apply for a [N,H,W,C] input       N*(H/2)*(W/2)*(C/8) uint16
 *   UNETK_WS_WGRAD        (mode, Cu, Cs) dw of unetk_wgrad                                           Cu*taps*Cs floats
 *   UNETK_WS_DICE_ACCUM   (C)            unetk_dice_ce_args.accum                                    3C+2 doubles
 *   UNETK_WS_DICE_COEF    (C)            unetk_dice_ce_args.coef                                     2C+1 floats
 *   UNETK_WS_EVAL_ACCUM   (N, C)         unetk_eval_args.accum                                       N*(3C+2) doubles
 *   UNETK_WS_CONFUSION    (C)            counts of unetk_argmax_confusion / unetk_eval_loss_metrics  4*C int64      */
#define UNETK_WS_BN_STATS 0
#define UNETK_WS_BN_BWD_SUMS 1
#define UNETK_WS_HEAD_BN_SUMS 2
#define UNETK_WS_POOL_IDX 3
#define UNETK_WS_WGRAD 4
#define UNETK_WS_DICE_ACCUM 5
#define UNETK_WS_DICE_COEF 6
#define UNETK_WS_EVAL_ACCUM 7
#define UNETK_WS_CONFUSION 8
UNETK_API int64_t unetk_query_workspace(int32_t what, int32_t a, int32_t b, int32_t c, int32_t d);
/* sizeof() of the argument structs as compiled into the library, for bindings that mirror them by hand (ctypes, cgo, JNI):
 * 0 unetk_tensor, 1 unetk_conv_args, 2 unetk_wgrad_args, 3 unetk_bn_finalize_args, 4 unetk_bn_bwd_args, 5 unetk_wjob,
 * 6 unetk_dice_ce_args, 7 unetk_head_bn_bwd_args, 8 unetk_eval_image, 9 unetk_eval_args; -1 for an unknown id.   */
UNETK_API int32_t unetk_struct_size(int32_t which);

/* ---- layout --------------------------------------------------------------------------------
 * unetk_im2col3x3_first: X.to(device) + first conv's implicit im2col (utils/training.py:45,
 * unet/unet.py:16 for down1).  x is NCHW fp32 [N,Cin,H,W]; out is NHWC [N,H,W,out.c] with
 * out[.., (r*3+s)*Cin + ci] = x[n, ci, h+r-1, w+s-1] (zero outside / for k >= 9*Cin).
 * The first conv then runs as a 1x1 contraction over out.c channels.                          */
UNETK_API int unetk_im2col3x3_first(const float* x_nchw, int32_t n, int32_t cin, int32_t h, int32_t w,
                          const unetk_tensor* out, void* stream);

/* unetk_permute3: dst[i0*ds0+i1*ds1+i2*ds2] = (dst_dtype) src[i0*ss0+i1*ss1+i2*ss2], fp32 source.
 * Packs nn.Conv2d / nn.ConvTranspose2d weights (OIHW / IOHW fp32, unet/unet.py:16,19,59) into the
 * K-major operand layouts of the contraction kernels and unpacks weight gradients back.       */
UNETK_API int unetk_permute3(const float* src, void* dst, int32_t dst_dtype, int32_t d0, int32_t d1, int32_t d2,
                   int64_t ss0, int64_t ss1, int64_t ss2, int64_t ds0, int64_t ds1, int64_t ds2,
                   void* stream);

/* Batched weight (un)packing: ONE launch for all layers instead of one unetk_permute3 per tensor.
 *   kind 0  conv3x3    param [co][ci][3][3]   <->  fwd pack [co][t][ci]        and  dgrad pack [ci][8-t][co]
 *   kind 1  convT2x2   param [ci][co][2][2]   <->  fwd pack [(a,b,co)][ci]     and  dgrad pack [ci][(a,b)][co]
 *   kind 2  first conv param [co][ci][3][3]   <->  fwd pack [co][kpad] with k = t*ci_count + ci (zero padded)
 * pack  : src = fp32 parameter, dst0 = fwd pack, dst1 = dgrad pack (NULL for kind 2), packs have dtype `dtype`.
 * unpack: src = fp32 weight-gradient workspace in the layout unetk_wgrad produces
 *         (kind 0: [co][t][ci]; kind 1: [ci][(a,b)][co]; kind 2: [co][kpad]), dst0 = fp32 gradient in parameter layout.
 * `jobs` and `tiles` are DEVICE arrays built once by the host; tiles[i] = {job, a0, b0, 0} enumerates the 32x32
 * (a,b) channel tiles of every job (kind 0: a = co, b = ci; kind 1: a = ci, b = co; kind 2: one tile {job,0,0,0}).
 * unpack: when `dst_base` is not NULL each job's dst0 is a BYTE OFFSET into dst_base (the gradient buffer is a fresh
 * allocation every backward pass, the job table is not). */
typedef struct unetk_wjob {
  const void* src;
  void* dst0;
  void* dst1;
  int32_t kind, cout, cin, kpad;
} unetk_wjob;
UNETK_API int unetk_weights_pack(const unetk_wjob* jobs, const int32_t* tiles, int32_t ntiles, int32_t dtype, void* stream);
UNETK_API int unetk_weights_unpack(const unetk_wjob* jobs, const int32_t* tiles, int32_t ntiles, void* dst_base, void* stream);

/* ---- contractions --------------------------------------------------------------------------
 * One implicit-GEMM entry for every "activation x weight" product of the path:
 *   mode 0  1x1           y[p, co]        = sum_ci          x[p, ci]            w[co][ci]
 *   mode 1  3x3 pad 1     y[p, co]        = sum_{t,ci}      x[p + off(t), ci]   w[co][t][ci]
 *           (nn.Conv2d k3 p1 forward, unet/unet.py:16,19; and its data gradient when w holds the
 *            flipped/transposed pack)
 *   mode 2  convT 2x2 s2  y[2p+(a,b), co] = sum_ci          x[p, ci]            w[(a,b,co)][ci]   (+bias)
 *           (nn.ConvTranspose2d forward, unet/unet.py:59; y may be a channel slice of the concat buffer,
 *            which replaces torch.cat at unet/unet.py:63)
 *   mode 3  convT gather  y[p, ci]        = sum_{(a,b),co}  x[2p+(a,b), co]     w[ci][(a,b)][co]
 *           (data gradient of ConvTranspose2d)
 * Optional: bias[C_out]; stat_sum/stat_sumsq[C_out] (double, ACCUMULATED) = per-channel sum and
 * sum of squares of y as stored (BatchNorm batch statistics, unet/unet.py:17,20).
 * Optional (data-gradient launches): when y is the gradient dA w.r.t. the ACTIVATED output of a conv+BN+ReLU layer,
 * bn_z (that layer's raw conv output, same shape as y) + bn_scale/bn_shift/bn_mean/bn_invstd make the kernel also
 * accumulate that layer's BatchNorm-backward reductions into bn_sums[2][C] (double):
 *   dy = dA * [relu(z*scale+shift) > 0];  bn_sums[0][c] += sum dy;  bn_sums[1][c] += sum dy * (z-mean)*invstd
 * i.e. unetk_bn_relu_bwd_reduce without a separate pass over dA.                                             */
typedef struct unetk_conv_args {
  unetk_tensor x;
  const void* w; /* same dtype as x */
  unetk_tensor y;
  int32_t mode;
  int32_t algo;
  const float* bias;
  double* stat_sum;
  double* stat_sumsq;
  unetk_tensor bn_z; /* ptr NULL = no fused BatchNorm-backward reduction */
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_invstd;
  double* bn_sums;
} unetk_conv_args;
UNETK_API int unetk_conv(const unetk_conv_args* a, void* stream);

/* Weight gradient: dw[cu][t][cs] += sum_p u[p, cu] * s[gather(p, t), cs]   (fp32, accumulated)
 *   mode 0: 1 tap; mode 1: 3x3 pad 1 (u = dY, s = X); mode 2: 2x2 stride 2 (u = X low-res, s = dY hi-res)
 * Replaces the filter-gradient half of convolution_backward for unet/unet.py:16,19,59.        */
typedef struct unetk_wgrad_args {
  unetk_tensor u;
  unetk_tensor s;
  float* dw;
  int32_t mode;
  int32_t algo;
  float* partial;        /* UNETK_TC_DETERMINISTIC: scratch of unetk_wgrad_partial_bytes(a) bytes (else NULL) */
  int64_t partial_bytes; /* size of `partial` */
} unetk_wgrad_args;
UNETK_API int unetk_wgrad(const unetk_wgrad_args* a, void* stream);
/* Bytes of `partial` scratch unetk_wgrad needs for this problem under UNETK_TC_DETERMINISTIC (0: none -- one split, or the
 * CUDA-core tier); depends on the device's SM count.  Negative on bad arguments. */
UNETK_API int64_t unetk_wgrad_partial_bytes(const unetk_wgrad_args* a);

/* per-channel sum over all pixels: out[c] += sum_p t[p, c]  (bias gradients of ConvTranspose2d) */
UNETK_API int unetk_channel_sum(const unetk_tensor* t, float* out, void* stream);
/* The same with a run-to-run reproducible result: block partial sums go to `scratch` (at least
 * UNETK_CHANNEL_SUM_BLOCKS * C floats) and are added in index order instead of with fp32 atomics. */
#define UNETK_CHANNEL_SUM_BLOCKS 512
UNETK_API int unetk_channel_sum_ordered(const unetk_tensor* t, float* out, float* scratch, int64_t scratch_bytes, void* stream);

/* ---- BatchNorm + ReLU + MaxPool (unet/unet.py:17-18,20-21,40) ------------------------------- */
UNETK_API int unetk_bn_stats(const unetk_tensor* z, double* sum, double* sumsq, void* stream);

typedef struct unetk_bn_finalize_args {
  const double* sum;
  const double* sumsq;
  int64_t count; /* N*H*W */
  int32_t c;
  int32_t training; /* 1: batch statistics + running-stat update; 0: running statistics */
  const float* gamma;
  const float* beta;
  const float* conv_bias; /* bias of the preceding conv: folded into mean / running_mean */
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  float momentum, eps;
  float* scale;  /* gamma * invstd */
  float* shift;  /* beta - mean*scale (training) */
  float* mean;   /* batch mean of the bias-free conv output */
  float* invstd;
} unetk_bn_finalize_args;
UNETK_API int unetk_bn_finalize(const unetk_bn_finalize_args* a, void* stream);

/* a = relu(z*scale+shift); optionally pooled = maxpool2x2(a) in the same pass (pooled->ptr may be NULL).
 * pool_idx (optional, [N,H/2,W/2,C/8] uint16): 2 bits per channel = window position (2*dy+dx) of the FIRST maximum in
 * scan order -- what torch's max_pool2d backward routes the gradient to; consumed by unetk_bn_relu_bwd_*.            */
UNETK_API int unetk_bn_relu_apply(const unetk_tensor* z, const float* scale, const float* shift,
                        const unetk_tensor* a, const unetk_tensor* pooled, uint16_t* pool_idx, void* stream);

/* Backward of (BN train -> ReLU [-> MaxPool]):
 *   dy = dA * [a > 0],  dA = dy_full (optional) + route(dpool) (optional; to the position stored in pool_idx)
 *   reduce: sums[0][c] += sum dy, sums[1][c] += sum dy * xhat
 *   apply : dz = scale * (dy - s1/M - xhat * s2/M);  dgamma = s2, dbeta = s1                   */
typedef struct unetk_bn_bwd_args {
  unetk_tensor z;
  unetk_tensor dy;    /* ptr NULL if absent */
  unetk_tensor dpool; /* ptr NULL if absent; [N,H/2,W/2,C] */
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  double* sums; /* [2][C] */
  unetk_tensor dz;
  float* dgamma;
  float* dbeta;
  const void* pool_idx; /* uint16 [N,H/2,W/2,C/8] from unetk_bn_relu_apply; required when dpool is given */
} unetk_bn_bwd_args;
UNETK_API int unetk_bn_relu_bwd_reduce(const unetk_bn_bwd_args* a, void* stream);
UNETK_API int unetk_bn_relu_bwd_apply(const unetk_bn_bwd_args* a, void* stream);

/* ---- 1x1 classifier head (unet/unet.py:91) -------------------------------------------------- */
/* logits NCHW fp32 [N,dout,H,W] = a[N,H,W,64] . w[dout][cin] + b */
UNETK_API int unetk_head_fprop(const unetk_tensor* a, const float* w, const float* b, int32_t dout,
                     float* logits_nchw, void* stream);
/* da = dlogits . w ; dw[dout][cin] += ; db[dout] += */
UNETK_API int unetk_head_bwd(const float* dlogits_nchw, const unetk_tensor* a, const float* w, int32_t dout,
                   const unetk_tensor* da, float* dw, float* db, void* stream);

/* ---- weighted Dice + CE loss (utils/weighted_loss.py:31-98,140-166) ---------------------------
 * input_kind selects what `logits` holds:
 *   UNETK_LOSS_LOGITS     class scores; softmax inside (WeightedDiceCELoss, utils/weighted_loss.py:102-166)
 *   UNETK_LOSS_PROBS_LOG  probabilities (apply_softmax=False); the likelihood term is NLLLoss(log(p + nll_eps))
 *                         (WeightedDiceNLLLoss with nll_nonlin = log(x + 1e-9), utils/weighted_loss.py:276-343,
 *                         prompt_based/prompt.ipynb:68-70)
 *   UNETK_LOSS_PROBS_RAW  probabilities; NLLLoss on the raw input (nll_nonlin=None)
 * The backward pass then returns d loss / d input for that kind of input. */
#define UNETK_LOSS_LOGITS 0
#define UNETK_LOSS_PROBS_LOG 1
#define UNETK_LOSS_PROBS_RAW 2
typedef struct unetk_dice_ce_args {
  const float* logits;   /* NCHW fp32 */
  const int64_t* target; /* [N,H,W] */
  int32_t n, c, h, w;
  const float* class_weights; /* NULL or [C] */
  int32_t has_ignore;
  int64_t ignore_index;
  float dice_weight, ce_weight, smooth;
  double* accum; /* [3C+2] zeroed by caller: I_c, P_c, G_c, ce_num, ce_den */
  float* coef;   /* [2C+1] written by fwd, read by bwd */
  float* loss;   /* [1] */
  int32_t* status; /* |= 1 if a label is outside [0,C) (the reference raises from scatter_) */
  const float* grad_out; /* [1] device scalar (bwd) */
  float* dlogits;        /* NCHW fp32 (bwd) */
  int32_t input_kind;    /* UNETK_LOSS_* */
  float nll_eps;         /* UNETK_LOSS_PROBS_LOG only */
} unetk_dice_ce_args;
UNETK_API int unetk_dice_ce_fwd(const unetk_dice_ce_args* a, void* stream);
UNETK_API int unetk_dice_ce_bwd(const unetk_dice_ce_args* a, void* stream);

/* ---- confusion-count metrics (utils/MetricsHistory.py:65-86) ---------------------------------- */
/* pred NCHW fp32 [N,C,H,W] (N images at once), label [N,H,W] int64.
 * counts[4][C] (tp, fp, fn, tn; int64) are ACCUMULATED; argmax_out (optional) receives the hard mask. */
UNETK_API int unetk_argmax_confusion(const float* pred, const int64_t* label, int32_t n, int32_t c, int32_t h,
                           int32_t w, int64_t* counts, uint8_t* argmax_out, int32_t* status, void* stream);

/* BatchNorm apply + ReLU of the last block fused with the head forward (unet/unet.py:20-21 -> :91).  `a` may be NULL
 * (or a->ptr NULL): the activation is then not stored at all -- with unetk_head_bn_bwd_* nothing reads it again.
 * dout 1..4 and C == 64 (the U-Net head), else UNETK_ERR_UNSUPPORTED (use unetk_bn_relu_apply + unetk_head_fprop). */
UNETK_API int unetk_bn_relu_head_fprop(const unetk_tensor* z, const float* scale, const float* shift, const unetk_tensor* a,
                             const float* w_head, const float* b_head, int32_t dout, float* logits_nchw, void* stream);

/* ---- head backward fused with the BatchNorm backward of the last block (unet/unet.py:91 <- :21-25) ----------- */
/* z is the raw conv output of the layer whose activation a = relu(z*scale+shift) feeds the 1x1 head.  Both passes
 * recompute da = dlogits . w_head per pixel, so the [N,H,W,C] activation gradient never touches memory.
 * sums: [(3+dout)*C] doubles zeroed by the caller: sum dy | sum dy*xhat | dW_head[k][c] (dout rows) | db_head[k] in the
 * first dout entries of the last row.
 * dout 1..4 (UNETK_ERR_UNSUPPORTED otherwise: use unetk_head_bwd + unetk_bn_relu_bwd_*). */
typedef struct unetk_head_bn_bwd_args {
  unetk_tensor z;
  const float* dlogits; /* NCHW fp32 [N,dout,H,W] */
  const float* w_head;  /* [dout][C] */
  int32_t dout;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  double* sums;
  unetk_tensor dz; /* apply only */
  float* dgamma;   /* apply only, [C] (may be NULL) */
  float* dbeta;
  float* dw_head;  /* apply only, [dout][C] */
  float* db_head;  /* apply only, [dout] (may be NULL) */
} unetk_head_bn_bwd_args;
UNETK_API int unetk_head_bn_bwd_reduce(const unetk_head_bn_bwd_args* a, void* stream);
UNETK_API int unetk_head_bn_bwd_apply(const unetk_head_bn_bwd_args* a, void* stream);

/* ---- evaluation tail (utils/utils.py:51-75,101-115; utils/training.py:93-101) ------------------- */
/* One entry per image of a batch (device array).  The network output [N,C,TH,TW] holds image i in the window
 * [crop_top, crop_top+crop_h) x [crop_left, crop_left+crop_w) (meta["pad"], meta["new_size"] of
 * resize_with_padding, utils/utils.py:42-47); (out_h, out_w) is meta["original_size"].  `offset` is the index of
 * the image's first pixel in the packed label buffer (and, times C, in the packed output of unetk_crop_resize,
 * where image i is stored as [C,out_h,out_w]). */
typedef struct unetk_eval_image {
  int32_t crop_top, crop_left, crop_h, crop_w;
  int32_t out_h, out_w;
  int64_t offset;
} unetk_eval_image;

/* process_batch_reverse (utils/utils.py:101-115): crop + F.interpolate(mode 0 'bilinear', align_corners=False |
 * mode 1 'nearest') of every image of the batch in one launch.  max_out_pixels = max_i out_h*out_w. */
UNETK_API int unetk_crop_resize(const float* src_nchw, int32_t n, int32_t c, int32_t th, int32_t tw,
                      const unetk_eval_image* images, int32_t max_out_pixels, int32_t mode, float* out_packed,
                      void* stream);

/* eval_loop body (utils/training.py:93-101) for a whole batch: per image, the bilinear-resized logits are reduced
 * to the Dice+CE loss of WeightedDiceCELoss (batch of one, as the reference calls it) and to argmax confusion
 * counts, without materialising the resized logits. */
typedef struct unetk_eval_args {
  const float* logits; /* NCHW fp32 [N,C,TH,TW] */
  int32_t n, c, th, tw;
  const unetk_eval_image* images; /* device [N] */
  int32_t max_out_pixels;
  const void* labels;  /* packed, image i at element images[i].offset */
  int32_t label_dtype; /* UNETK_U8 | UNETK_I64 */
  const float* class_weights; /* NULL or [C] */
  int32_t has_ignore;
  int64_t ignore_index;
  float dice_weight, ce_weight, smooth;
  double* accum;         /* [N][3C+2] zeroed by the caller */
  float* loss_per_image; /* [N] written */
  double* loss_sum;      /* NULL or [1]: += sum_i (double)loss_i, in image order (total_loss += loss.item()) */
  int64_t* counts;       /* [4][C] tp, fp, fn, tn ACCUMULATED (MetricsHistory.accumulate) */
  int32_t* status;       /* |= 1 if a label is outside [0,C) */
} unetk_eval_args;
UNETK_API int unetk_eval_loss_metrics(const unetk_eval_args* a, void* stream);

/* ---- other model families on the same blocks ------------------------------------ */
/* Reconstruction output (autoencoder/autoencoder.py:188-191): out[n,k,h,w] = sigmoid(z[n,h,w,k] + bias[k]) for k < dout
 * (z is the NHWC output of the 3x3 convolution whose Cout is zero-padded; only its first dout channels are read).
 * Backward: dz[p,k] = dy[p,k] * out (1 - out) for k < dout and 0 for k in [dout, 8); channels >= 8 of dz are NOT written
 * (the caller zeroes dz once); dbias[k] += sum_p dz[p,k] (may be NULL).  dout 1..8.                                  */
UNETK_API int unetk_bias_sigmoid_fwd(const unetk_tensor* z, const float* bias, int32_t dout, float* out_nchw, void* stream);
UNETK_API int unetk_bias_sigmoid_bwd(const float* dy_nchw, const float* out_nchw, int32_t dout, const unetk_tensor* dz,
                                     float* dbias, void* stream);

/* F.interpolate(src, size=(dst.h, dst.w), mode='bilinear', align_corners=False) in NHWC (clip/clipunet.py:99-100); dst may
 * be a channel slice of a concat buffer.  _bwd computes d src from d dst (gather over the interpolation footprints). */
UNETK_API int unetk_bilinear_up_fwd(const unetk_tensor* src, const unetk_tensor* dst, void* stream);
UNETK_API int unetk_bilinear_up_bwd(const unetk_tensor* ddst, const unetk_tensor* dsrc, void* stream);

/* PromptModel.forward tail (prompt_based/prompt.py:36-56), NCHW fp32: p = softmax(clip_logits [N,4,H,W]),
 * m = sigmoid(mask_logits [N,1,H,W]); final = [1 - m, m p0 + m p3, m p1, m p2].  The CLIP branch is frozen in the
 * reference (prompt.py:30-31), so the backward returns d mask_logits only.                                         */
UNETK_API int unetk_prompt_compose_fwd(const float* clip_logits, const float* mask_logits, int32_t n, int32_t h, int32_t w,
                                       float* final_probs, void* stream);
UNETK_API int unetk_prompt_compose_bwd(const float* clip_logits, const float* mask_logits, const float* dfinal, int32_t n,
                                       int32_t h, int32_t w, float* dmask_logits, void* stream);

/* Layout converters of the block-level API (DoubleConvReLU / Down / Up called on their own, unet/unet.py:24,44,62):
 * NCHW fp32 [N,C,H,W] <-> NHWC dst/src (dtype of the tensor descriptor; any C). */
UNETK_API int unetk_nchw_to_nhwc(const float* src_nchw, const unetk_tensor* dst, void* stream);
UNETK_API int unetk_nhwc_to_nchw(const unetk_tensor* src, float* dst_nchw, void* stream);

/* ---- data-parallel exchange step ---------------------------------------------------------
 * In-place sum all-reduce of a fp32 buffer that every rank holds in SYMMETRIC memory bound to one NVSwitch multicast
 * object (NVLS): rank r reduces and re-broadcasts the r-th slice with multimem.ld_reduce / multimem.st; `scale` is applied
 * to the sum (1/world for a mean).  `multicast_ptr` is the multicast address of the buffer (16-byte aligned, n_elems a
 * multiple of 4).  The CALLER provides the cross-rank barriers: one before the launch (every rank's buffer is written and
 * visible system-wide) and one after it (every slice has been stored on every rank).                                    */
UNETK_API int unetk_nvls_allreduce_f32(void* multicast_ptr, int64_t n_elems, int32_t rank, int32_t world, float scale,
                                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETK_H */
