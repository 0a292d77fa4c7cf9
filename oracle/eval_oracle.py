"""CPU restatement of the reference's evaluation tail (TEST INFRASTRUCTURE ONLY -- never imported by the product).

What it restates (reference file:line):
  * utils/utils.py:51-75   ``reverse_resize_and_padding``: crop ``[top:top+new_h, left:left+new_w]`` out of the
                           network output and ``F.interpolate(..., mode='bilinear', align_corners=False)`` (or
                           ``'nearest'``) back to the original size
  * utils/utils.py:13-49   ``resize_with_padding`` metadata arithmetic (scale, rounded new size, centred padding)
  * utils/training.py:89-101  per image: ``loss_fn(pred[None], label[None])``, ``total_loss += loss.item()``,
                           ``agg.accumulate(pred, label)``; ``avg_loss = total_loss / num_images``

The interpolation arithmetic lives in PyTorch (ATen ``upsample_bilinear2d``, not vendored in the reference); its
published algorithm is restated here in numpy float32 with one rounding per operation, in this order:

    scale = float32(in) / float32(out)
    src   = max(fma(scale, dst + 0.5, -0.5), 0)         i0 = int(src)   i1 = i0 + (i0 < in - 1)
    l1    = src - i0                                    l0 = 1 - l1
    out   = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11)

The CUDA kernel follows exactly this operation order with round-to-nearest, non-contracted float ops, so it is
compared bit-for-bit against this file; this file is pinned against the unmodified reference through
tests/golden/eval.npz to 2e-6 absolute (ATen's CPU build contracts the remaining multiply-adds into FMAs in a
vector-width dependent order, so the last bit of about half the elements differs; with a non-fused ``src`` the
difference would be 1.4e-5, which is how the fused form was identified).
"""
from __future__ import annotations

import numpy as np
import torch

from . import loss_oracle, metrics_oracle

F32 = np.float32


def resize_meta(orig_h: int, orig_w: int, target_size: int):
    """Metadata of ``resize_with_padding`` (utils/utils.py:25-48) without touching pixels."""
    scale = min(target_size / orig_w, target_size / orig_h)
    new_w, new_h = int(round(orig_w * scale)), int(round(orig_h * scale))
    pad_w, pad_h = target_size - new_w, target_size - new_h
    left, top = pad_w // 2, pad_h // 2
    return {"original_size": (orig_h, orig_w), "new_size": (new_h, new_w),
            "pad": (left, top, pad_w - left, pad_h - top), "scale": scale}


def _axis_bilinear(n_in: int, n_out: int):
    scale = F32(n_in) / F32(n_out)
    dst = np.arange(n_out, dtype=F32)
    # fma(scale, dst + 0.5, -0.5): the float64 product of two float32 numbers and the subtraction of 0.5 are both
    # exact, so the single rounding to float32 below is the fused result
    src = (np.float64(scale) * (dst + F32(0.5)).astype(np.float64) - 0.5).astype(F32)
    src = np.maximum(src, F32(0.0)).astype(F32)
    i0 = src.astype(np.int64)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (src - i0.astype(F32)).astype(F32)
    l0 = (F32(1.0) - l1).astype(F32)
    return i0, i1, l0, l1


def _axis_nearest(n_in: int, n_out: int):
    scale = F32(n_in) / F32(n_out)
    idx = np.floor(np.arange(n_out, dtype=F32) * scale).astype(np.int64)
    return np.minimum(idx, n_in - 1)


def crop_resize(image: np.ndarray, meta: dict, interpolation: str = "bilinear") -> np.ndarray:
    """(C,T,T) float32 -> (C,orig_h,orig_w) float32, utils/utils.py:62-75."""
    image = np.asarray(image, dtype=F32)
    left, top, _, _ = meta["pad"]
    new_h, new_w = meta["new_size"]
    oh, ow = meta["original_size"]
    crop = image[..., top: top + new_h, left: left + new_w]
    if interpolation == "nearest":
        return np.ascontiguousarray(crop[..., _axis_nearest(new_h, oh)[:, None], _axis_nearest(new_w, ow)[None, :]])
    if interpolation != "bilinear":
        raise ValueError(interpolation)
    y0, y1, ly0, ly1 = _axis_bilinear(new_h, oh)
    x0, x1, lx0, lx1 = _axis_bilinear(new_w, ow)
    v00 = crop[..., y0[:, None], x0[None, :]]
    v01 = crop[..., y0[:, None], x1[None, :]]
    v10 = crop[..., y1[:, None], x0[None, :]]
    v11 = crop[..., y1[:, None], x1[None, :]]
    lx0, lx1 = lx0[None, :], lx1[None, :]
    ly0, ly1 = ly0[:, None], ly1[:, None]
    top_row = (lx0 * v00).astype(F32) + (lx1 * v01).astype(F32)
    bot_row = (lx0 * v10).astype(F32) + (lx1 * v11).astype(F32)
    return ((ly0 * top_row.astype(F32)).astype(F32) + (ly1 * bot_row.astype(F32)).astype(F32)).astype(F32)


def eval_batch(logits: np.ndarray, metas, labels, num_classes: int, loss_kwargs: dict, interpolation="bilinear"):
    """One batch of utils/training.py:93-101.  Returns (per-image float32 losses, int64 counts [4,C], resized preds)."""
    losses, preds = [], []
    counts = np.zeros((4, num_classes), dtype=np.int64)
    for img, meta, label in zip(logits, metas, labels):
        pred = crop_resize(img, meta, interpolation)
        label = np.asarray(label).reshape(pred.shape[1:]).astype(np.int64)
        loss = loss_oracle.dice_ce_loss(torch.from_numpy(pred)[None], torch.from_numpy(label)[None],
                                        dtype=torch.float32, **loss_kwargs)
        losses.append(np.float32(loss.item()))
        counts += np.stack(metrics_oracle.confusion_counts(pred, label, num_classes))
        preds.append(pred)
    return np.array(losses, dtype=F32), counts, preds


def eval_epoch(batches, num_classes: int, loss_kwargs: dict, ignore_index=None):
    """``batches`` = iterable of (logits [N,C,T,T], metas, labels).  Returns what ``eval_loop`` returns
    (avg_loss, mean_dice, mean_iou) plus the counts (utils/training.py:103-121)."""
    total, n_img = 0.0, 0
    counts = np.zeros((4, num_classes), dtype=np.int64)
    for logits, metas, labels in batches:
        losses, c, _ = eval_batch(logits, metas, labels, num_classes, loss_kwargs)
        for v in losses:
            total += float(v)           # total_loss += loss.item()
        n_img += len(losses)
        counts += c
    mean_dice, mean_iou, _, _, _, _ = metrics_oracle.epoch_metrics(*counts, ignore_index=ignore_index)
    return total / n_img, mean_dice, mean_iou, counts
