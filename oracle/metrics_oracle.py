"""CPU oracle for the confusion-count metrics (TEST INFRASTRUCTURE ONLY; numpy, integer-exact).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this file; the product never does.

Restates ``utils/MetricsHistory.py`` (citations into ``/root/reference``):

* ``accumulate``             utils/MetricsHistory.py:55-86   argmax over classes (ties -> lowest
  index, NaN counts as the maximum like torch.argmax), per-class TP/FP/FN/TN via a confusion
  matrix: cm = bincount(label*C + pred); tp = diag; fp = colsum - tp; fn = rowsum - tp;
  tn = total - tp - fp - fn.
* ``compute_epoch_metrics``  utils/MetricsHistory.py:89-128  iou = tp/(tp+fp+fn),
  dice = 2tp/(2tp+fp+fn), acc = (tp+tn)/(tp+tn+fp+fn); macro mean over classes except
  ``ignore_index`` (a CLASS is dropped from the mean, pixels are never dropped).
"""
from __future__ import annotations

import numpy as np


def argmax_first(pred: np.ndarray) -> np.ndarray:
    """argmax over axis 0 of [C,H,W]; first maximum wins; NaN is treated as the largest value."""
    c = pred.shape[0]
    best = pred[0].copy()
    idx = np.zeros(pred.shape[1:], dtype=np.int64)
    for k in range(1, c):
        v = pred[k]
        take = (v > best) | (np.isnan(v) & ~np.isnan(best))
        best = np.where(take, v, best)
        idx = np.where(take, k, idx)
    return idx


def confusion_counts(pred: np.ndarray, label: np.ndarray, num_classes: int):
    """pred [C,H,W] float, label [H,W] or [1,H,W] int -> (tp, fp, fn, tn) int64 arrays of length C."""
    label = np.asarray(label).reshape(pred.shape[1:]).astype(np.int64)
    if label.min(initial=0) < 0 or label.max(initial=0) >= num_classes:
        raise RuntimeError("label out of range for one_hot")
    hard = argmax_first(np.asarray(pred))
    cm = np.bincount((label * num_classes + hard).ravel(), minlength=num_classes * num_classes)
    cm = cm.reshape(num_classes, num_classes).astype(np.int64)
    tp = np.diag(cm).copy()
    fp = cm.sum(axis=0) - tp
    fn = cm.sum(axis=1) - tp
    tn = cm.sum() - tp - fp - fn
    return tp, fp, fn, tn


def epoch_metrics(tp, fp, fn, tn, ignore_index=None):
    """(mean_dice, mean_iou, mean_acc, per_class_dice, per_class_iou, per_class_acc) in float64."""
    tp, fp, fn, tn = (np.asarray(a, dtype=np.float64) for a in (tp, fp, fn, tn))
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = tp / (tp + fp + fn)
        dice = 2 * tp / (2 * tp + fp + fn)
        acc = (tp + tn) / (tp + tn + fp + fn)
    mask = np.ones(len(tp), dtype=bool)
    if ignore_index is not None and 0 <= ignore_index < len(tp):
        mask[ignore_index] = False
    return dice[mask].mean(), iou[mask].mean(), acc[mask].mean(), dice, iou, acc
