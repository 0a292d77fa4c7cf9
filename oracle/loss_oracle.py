"""CPU oracle for the weighted Dice + cross-entropy loss (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this file; the product never does.

Closed-form restatement of ``utils/weighted_loss.py`` (citations into ``/root/reference``):

* ``WeightedMemoryEfficientDiceLoss.forward``  utils/weighted_loss.py:31-98
    p = softmax(x, 1)                                   (:36)
    I_c = sum_{n,h,w} p_c [y == c];  P_c = sum p_c;  G_c = #{y == c}   (:57-73, batch AND pixels)
    dc_c = (2 I_c + s) / clip(P_c + G_c + s, 1e-8)      (:76-77)
    class ``ignore_index`` is dropped from the mean only (:79-85) -- pixels are never masked
    (the ``mask`` variable is always None, :49)
    result = -( sum_c dc_c w_c / clamp(sum_c w_c, 1e-8) )  or  -mean_c dc_c   (:87-98)
* ``WeightedDiceCELoss.forward``               utils/weighted_loss.py:140-166
    dice_weight * dice + ce_weight * CrossEntropyLoss(weight=class_weights, ignore_index=..)
    CE = sum_valid w[y] * nll / sum_valid w[y]  (torch semantics; plain mean without weights)

Everything is evaluated in float64 unless ``dtype`` says otherwise; gradients come from
autograd over this closed form, plus an explicit analytic gradient (``dice_ce_grad``) that the
CUDA backward kernel mirrors.
"""
from __future__ import annotations

from typing import Optional

import torch


def _onehot(p, target):
    """utils/weighted_loss.py:50-58: when probs.shape == y.shape (i.e. C == 1 with [N,1,H,W] targets)
    the reference uses ``y.float()`` itself as the one-hot; otherwise zeros_like + scatter_."""
    if p.shape[1] == 1:
        return target.unsqueeze(1).to(p.dtype)
    return torch.zeros_like(p).scatter_(1, target.unsqueeze(1), 1.0)


def dice_ce_terms(logits: torch.Tensor, target: torch.Tensor, num_classes: int):
    """Per-class sums I, P, G and per-pixel log-probabilities.  target: [N,H,W] int64."""
    logp = torch.log_softmax(logits, dim=1)
    p = logp.exp()
    onehot = _onehot(p, target)
    inter = (p * onehot).sum(dim=(0, 2, 3))
    psum = p.sum(dim=(0, 2, 3))
    gsum = onehot.sum(dim=(0, 2, 3))
    return inter, psum, gsum, logp


def dice_ce_loss(logits: torch.Tensor, target: torch.Tensor, *, dice_weight: float = 1.0,
                 ce_weight: float = 1.0, ignore_index: Optional[int] = None,
                 class_weights: Optional[torch.Tensor] = None, smooth_dice: float = 1e-5,
                 dtype=torch.float64) -> torch.Tensor:
    """Scalar loss of WeightedDiceCELoss(...)(logits, target).  target: [N,H,W] or [N,1,H,W]."""
    if target.ndim == 4:
        if target.shape[1] != 1:
            raise ValueError("target must be [N,H,W] or [N,1,H,W]")
        target = target[:, 0]
    elif target.ndim != 3:
        raise ValueError("target must be [N,H,W] or [N,1,H,W]")
    target = target.long()
    x = logits.to(dtype)
    c = x.shape[1]
    inter, psum, gsum, logp = dice_ce_terms(x, target, c)
    dc = (2.0 * inter + smooth_dice) / torch.clip(psum + gsum + smooth_dice, 1e-8)
    valid = torch.ones(c, dtype=torch.bool)
    if ignore_index is not None and 0 <= ignore_index < c:
        valid[ignore_index] = False
    if class_weights is not None:
        w = class_weights.to(dtype)
        dice = (dc[valid] * w[valid]).sum() / w[valid].sum().clamp(min=1e-8)
    else:
        dice = dc[valid].mean()
    nll = -logp.gather(1, target.unsqueeze(1))[:, 0]
    pix_valid = torch.ones_like(target, dtype=torch.bool) if ignore_index is None else (target != ignore_index)
    if class_weights is not None:
        wy = class_weights.to(dtype)[target] * pix_valid
    else:
        wy = pix_valid.to(dtype)
    ce = (wy * nll).sum() / wy.sum()
    return dice_weight * (-dice) + ce_weight * ce


def dice_ce_grad(logits: torch.Tensor, target: torch.Tensor, *, dice_weight: float = 1.0,
                 ce_weight: float = 1.0, ignore_index: Optional[int] = None,
                 class_weights: Optional[torch.Tensor] = None, smooth_dice: float = 1e-5,
                 dtype=torch.float64) -> torch.Tensor:
    """Analytic d loss / d logits (the formula the CUDA backward kernel implements).

    With D_c = P_c + G_c + s (unclipped branch) and a_c = dice_weight * w_c / sum_valid w:
        dL/dp_c(pixel) = -a_c * (2 [y==c] - dc_c) / D_c            =: g_c   (0 for the ignored class)
        dL/dx_k = p_k (g_k - sum_c p_c g_c) + ce_weight * wy (p_k - [y==k]) / sum wy
    """
    if target.ndim == 4:
        target = target[:, 0]
    target = target.long()
    x = logits.to(dtype)
    c = x.shape[1]
    inter, psum, gsum, logp = dice_ce_terms(x, target, c)
    p = logp.exp()
    onehot = _onehot(p, target)
    den = psum + gsum + smooth_dice
    clipped = den < 1e-8
    den_c = torch.clip(den, 1e-8)
    dc = (2.0 * inter + smooth_dice) / den_c
    valid = torch.ones(c, dtype=torch.bool)
    if ignore_index is not None and 0 <= ignore_index < c:
        valid[ignore_index] = False
    if class_weights is not None:
        w = class_weights.to(dtype) * valid
        a = w / w.sum().clamp(min=1e-8)
    else:
        a = valid.to(dtype) / valid.sum()
    a = a * dice_weight
    # d dc_c / d p_c(pixel) = (2 [y==c] - dc_c [den not clipped]) / den_c
    g = -(a / den_c)[None, :, None, None] * (2.0 * onehot - (dc * (~clipped))[None, :, None, None])
    gx = p * (g - (p * g).sum(dim=1, keepdim=True))
    pix_valid = torch.ones_like(target, dtype=torch.bool) if ignore_index is None else (target != ignore_index)
    if class_weights is not None:
        wy = class_weights.to(dtype)[target] * pix_valid
    else:
        wy = pix_valid.to(dtype)
    ce_onehot = torch.zeros_like(p).scatter_(1, target.unsqueeze(1), 1.0)   # CE always uses the true one-hot
    gx = gx + ce_weight * (wy / wy.sum())[:, None] * (p - ce_onehot)
    return gx
