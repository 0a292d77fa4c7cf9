// Pieces shared by the training loss kernels (head_loss.cu) and the evaluation tail (eval.cu).
#pragma once
#include "common.cuh"

namespace unetk {

constexpr int kMaxClasses = 8;

// CM = compile-time bound on the class count (4 for the U-Net's 1/3/4 classes, kMaxClasses otherwise): the softmax
// costs CM exponentials per pixel, so the bound matters (the loss kernels are instruction-bound, not memory-bound)
template <int CM>
struct SoftmaxT {
  float p[CM];
  float logp_y;
};
using Softmax = SoftmaxT<kMaxClasses>;

// softmax over x[0..c) (entries >= c must be -inf) and log p[y]
template <int CM>
__device__ __forceinline__ void softmax_of(const float (&x)[CM], int c, int y, SoftmaxT<CM>& s) {
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < CM; ++k) m = fmaxf(m, x[k]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < CM; ++k) {
    s.p[k] = k < c ? expf(x[k] - m) : 0.f;
    sum += s.p[k];
  }
  const float inv = 1.f / sum;
  float xy = 0.f;
#pragma unroll
  for (int k = 0; k < CM; ++k) {
    s.p[k] *= inv;
    if (k == y) xy = x[k];
  }
  s.logp_y = xy - m - logf(sum);
}

template <int CM>
__device__ __forceinline__ void pixel_softmax(const float* __restrict__ logits, int64_t base, int64_t hw, int c, int y,
                                              SoftmaxT<CM>& s) {
  float x[CM];
#pragma unroll
  for (int k = 0; k < CM; ++k) x[k] = k < c ? logits[base + k * hw] : -INFINITY;
  softmax_of<CM>(x, c, y, s);
}

// Probability inputs (the prompt model's losses, utils/weighted_loss.py:207-210,338-340): p = x as given, and the
// "log-probability" of the label is nll_nonlin(x)[y] = log(x[y] + eps) (UNETK_LOSS_PROBS_LOG) or x[y] itself
// (UNETK_LOSS_PROBS_RAW: nll_nonlin=None feeds NLLLoss with the raw input).
template <int CM>
__device__ __forceinline__ void pixel_probs(const float* __restrict__ x, int64_t base, int64_t hw, int c, int y, int kind,
                                            float eps, SoftmaxT<CM>& s) {
  float py = 0.f;
#pragma unroll
  for (int k = 0; k < CM; ++k) {
    s.p[k] = k < c ? x[base + k * hw] : 0.f;
    if (k == y) py = s.p[k];
  }
  s.logp_y = kind == UNETK_LOSS_PROBS_LOG ? logf(py + eps) : py;
}

// torch.argmax over x[0..c): first maximum wins, NaN counts as the largest value (utils/MetricsHistory.py:65)
__device__ __forceinline__ int argmax_first(const float (&x)[kMaxClasses], int c) {
  float best = x[0];
  int arg = 0;
#pragma unroll
  for (int k = 1; k < kMaxClasses; ++k) {
    if (k < c && (x[k] > best || (isnan(x[k]) && !isnan(best)))) {
      best = x[k];
      arg = k;
    }
  }
  return arg;
}

// Loss from the per-class sums accum = [I_c | P_c | G_c | ce_num, ce_den] (utils/weighted_loss.py:76-98,165).
// coef (optional, [2C+1]) receives what the backward kernel needs.
__device__ inline float finalize_dice_ce(const double* accum, int c, const float* class_weights, int has_ignore,
                                         int64_t ignore_index, float smooth, float dice_weight, float ce_weight,
                                         float* coef) {
  double wsum = 0.0;
  for (int k = 0; k < c; ++k) {
    const bool valid = !(has_ignore && ignore_index >= 0 && ignore_index < c && k == ignore_index);
    if (valid) wsum += class_weights ? (double)class_weights[k] : 1.0;
  }
  if (class_weights && wsum < 1e-8) wsum = 1e-8;
  double dice = 0.0;
  for (int k = 0; k < c; ++k) {
    const bool valid = !(has_ignore && ignore_index >= 0 && ignore_index < c && k == ignore_index);
    const double I = accum[k], P = accum[c + k], G = accum[2 * c + k];
    const double den = P + G + (double)smooth;
    const double den_c = den < 1e-8 ? 1e-8 : den;
    const double dc = (2.0 * I + (double)smooth) / den_c;
    const double ak = valid ? (class_weights ? (double)class_weights[k] : 1.0) / wsum : 0.0;
    dice += ak * dc;
    if (coef) {
      coef[k] = (float)((double)dice_weight * ak / den_c);
      coef[c + k] = den < 1e-8 ? 0.f : (float)dc;
    }
  }
  const double ce_num = accum[3 * c], ce_den = accum[3 * c + 1];
  const double ce = ce_num / ce_den;  // NaN if every pixel is ignored, like torch
  if (coef) coef[2 * c] = (float)((double)ce_weight / ce_den);
  return (float)((double)dice_weight * (-dice) + (double)ce_weight * ce);
}

}  // namespace unetk
