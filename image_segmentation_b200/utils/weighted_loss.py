"""Drop-in for the reference's ``utils/weighted_loss.py``: ``WeightedMemoryEfficientDiceLoss`` (:6-98),
``WeightedDiceCELoss`` (:102-166) and the prompt-model variants ``WeightedMemoryEfficientDiceLossPrompt`` (:170-273)
and ``WeightedDiceNLLLoss`` (:276-343: Dice on PROBABILITIES + ``NLLLoss(log(p + 1e-9))``).

Same constructor arguments, same ``forward(outputs, targets)`` contract and error behaviour, but the
~12 ATen ops of the reference (softmax, zeros_like + scatter_, three [N,C,H,W] products/sums, clip,
log_softmax + nll_loss2d) are two CUDA kernels: a single reduction pass over the logits producing the
per-class Dice sums and the weighted CE sums, and a single elementwise backward pass producing
d loss / d logits.  Sums are accumulated in float64.

Out-of-range labels: the reference raises from ``scatter_`` synchronously.  Here the kernel sets a
device flag; it is checked without stalling the stream on the *next* call (or immediately when
``UNETK_STRICT_LABELS=1``), and raises the same ``RuntimeError``.
"""
import os
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib as L

_MAX_CLASSES = 8


class _Status:
    """Device flag mirrored to pinned host memory without synchronising the stream."""

    def __init__(self, device):
        self.dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.event = None

    def publish(self):
        self.host.copy_(self.dev, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def check(self, strict: bool, what: str):
        if self.event is None:
            return
        if strict:
            self.event.synchronize()
        if self.event.query() and int(self.host[0]) != 0:
            self.dev.zero_()
            self.host.zero_()
            self.event = None
            raise RuntimeError(f"{what}: label value outside [0, num_classes) "
                               "(the reference raises 'index out of bounds' from scatter_/one_hot)")


class _DiceCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, cfg):
        n, c, h, w = logits.shape
        dev = logits.device
        accum = torch.zeros(3 * c + 2, dtype=torch.float64, device=dev)
        coef = torch.empty(2 * c + 1, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        cw = cfg["class_weights"]
        args = L.DiceCeArgs(logits.data_ptr(), target.data_ptr(), n, c, h, w, L.ptr(cw),
                            1 if cfg["ignore_index"] is not None else 0,
                            cfg["ignore_index"] if cfg["ignore_index"] is not None else 0,
                            cfg["dice_weight"], cfg["ce_weight"], cfg["smooth"], accum.data_ptr(), coef.data_ptr(),
                            loss.data_ptr(), cfg["status"].dev.data_ptr(), None, None, cfg.get("input_kind", 0),
                            cfg.get("nll_eps", 0.0))
        L.dice_ce_fwd(args)
        ctx.save_for_backward(logits, target, coef)
        ctx.cfg = cfg
        ctx.keep = (cw,)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        logits, target, coef = ctx.saved_tensors
        cfg = ctx.cfg
        n, c, h, w = logits.shape
        dlogits = torch.empty_like(logits)
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        cw = cfg["class_weights"]
        args = L.DiceCeArgs(logits.data_ptr(), target.data_ptr(), n, c, h, w, L.ptr(cw),
                            1 if cfg["ignore_index"] is not None else 0,
                            cfg["ignore_index"] if cfg["ignore_index"] is not None else 0,
                            cfg["dice_weight"], cfg["ce_weight"], cfg["smooth"], None, coef.data_ptr(), None,
                            cfg["status"].dev.data_ptr(), go.data_ptr(), dlogits.data_ptr(), cfg.get("input_kind", 0),
                            cfg.get("nll_eps", 0.0))
        L.dice_ce_bwd(args)
        return dlogits, None, None


class _FusedLossBase(nn.Module):
    def _run(self, logits, target, dice_weight, ce_weight, smooth, ignore_index, class_weights, input_kind=L.LOSS_LOGITS,
             nll_eps=0.0):
        L.require_cuda(logits, target)
        if logits.dim() != 4:
            raise ValueError(f"expected logits [N,C,H,W], got {tuple(logits.shape)}")
        n, c, h, w = logits.shape
        if c > _MAX_CLASSES:
            raise NotImplementedError(f"fused Dice+CE kernel supports up to {_MAX_CLASSES} classes, got {c}")
        if target.shape != (n, h, w):
            raise ValueError(f"Shape mismatch: logits {tuple(logits.shape)}, target {tuple(target.shape)}")
        dev = logits.device
        if not hasattr(self, "_status") or self._status.dev.device != dev:
            self._status = _Status(dev)
        strict = os.environ.get("UNETK_STRICT_LABELS", "0") == "1"
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._status.check(False, type(self).__name__)
        cw = None
        if class_weights is not None:
            # device copy of the class weights, cached (a per-call H2D copy would stall the stream and break graph capture)
            key = (id(class_weights), class_weights._version, str(dev))
            if getattr(self, "_cw_key", None) != key:
                self._cw_dev = class_weights.detach().to(device=dev, dtype=torch.float32).contiguous()
                self._cw_key = key
            cw = self._cw_dev
            if cw.numel() != c:
                raise RuntimeError(f"weight tensor should be defined for all {c} classes, got {cw.numel()}")
        if ignore_index is not None and not (-2 ** 62 < int(ignore_index) < 2 ** 62):
            raise ValueError("ignore_index out of range")
        cfg = dict(class_weights=cw, ignore_index=None if ignore_index is None else int(ignore_index),
                   dice_weight=float(dice_weight), ce_weight=float(ce_weight), smooth=float(smooth),
                   status=self._status, input_kind=int(input_kind), nll_eps=float(nll_eps))
        lg = logits if (logits.dtype == torch.float32 and logits.is_contiguous()) else logits.float().contiguous()
        tg = target if (target.dtype == torch.int64 and target.is_contiguous()) else target.long().contiguous()
        with torch.cuda.device(dev):
            loss = _DiceCEFunction.apply(lg, tg, cfg)
            if not capturing:
                self._status.publish()
                if strict:
                    self._status.check(True, type(self).__name__)
        return loss


class WeightedMemoryEfficientDiceLoss(_FusedLossBase):
    """Soft Dice loss over the batch (reference: utils/weighted_loss.py:6-98); returns ``-dice``.

    Args:
        apply_softmax (bool): must be True (the U-Net path always passes logits).
        ignore_index (int, optional): class dropped from the class mean (pixels are never masked).
        class_weights (torch.Tensor, optional): per-class weights of the mean.
        smooth (float): Dice smoothing term.
    """

    def __init__(self, apply_softmax: bool = True, ignore_index: Optional[int] = None,
                 class_weights: Optional[torch.Tensor] = None, smooth: float = 1e-5):
        super().__init__()
        self.apply_softmax = apply_softmax
        self.ignore_index = ignore_index
        self.smooth = smooth
        self.class_weights = class_weights if class_weights is not None else None

    def forward(self, x, y):
        if not self.apply_softmax:
            raise NotImplementedError("apply_softmax=False (probability inputs) belongs to the prompt-model losses, "
                                      "which are outside the U-Net training path")
        if y.dim() == x.dim() and y.shape[1] == 1:
            y3 = y[:, 0]
        elif y.dim() == x.dim() and y.shape == x.shape:
            raise NotImplementedError("soft / one-hot float targets are not supported by the fused kernel")
        else:
            # the reference's [N,H,W] branch can never match (utils/weighted_loss.py:43) and raises too
            raise ValueError(f"Shape mismatch: probs {x.shape}, y {y.shape}")
        return self._run(x, y3, 1.0, 0.0, self.smooth, self.ignore_index, self.class_weights)


class WeightedDiceCELoss(_FusedLossBase):
    """``dice_weight * SoftDice + ce_weight * CrossEntropy`` (reference: utils/weighted_loss.py:102-166).

    Args:
        dice_weight (float), ce_weight (float): weights of the two terms.
        ignore_index (int, optional): class dropped from the Dice mean and pixel label ignored by CE.
        class_weights (torch.Tensor, optional): per-class weights for both terms.
        smooth_dice (float): Dice smoothing term.
        ce_kwargs (dict): extra CrossEntropyLoss arguments; only ``reduction='mean'`` is supported.
    """

    def __init__(self, dice_weight: float = 1.0, ce_weight: float = 1.0, ignore_index: Optional[int] = None,
                 class_weights: Optional[torch.Tensor] = None, smooth_dice: float = 1e-5, ce_kwargs={}):
        super().__init__()
        self.dice_weight = dice_weight
        self.ce_weight = ce_weight
        self.ignore_index = ignore_index
        self.class_weights = class_weights
        self.smooth_dice = smooth_dice
        extra = {k: v for k, v in dict(ce_kwargs).items() if not (k == "reduction" and v == "mean")}
        if "ignore_index" in extra and ignore_index is None:
            self._ce_ignore = extra.pop("ignore_index")
        else:
            extra.pop("ignore_index", None)
            self._ce_ignore = None
        if "weight" in extra and class_weights is None:
            raise NotImplementedError("pass class weights through class_weights=, not ce_kwargs['weight']")
        extra.pop("weight", None)
        if extra:
            raise NotImplementedError(f"ce_kwargs {sorted(extra)} are not supported by the fused Dice+CE kernel")
        if self._ce_ignore is not None:
            raise NotImplementedError("a CE-only ignore_index (via ce_kwargs) is not supported; use ignore_index=")
        # kept for API parity with the reference object graph
        self.dice = WeightedMemoryEfficientDiceLoss(True, ignore_index, class_weights, smooth_dice)

    def forward(self, outputs, targets):
        if targets.ndim == 3:
            t3 = targets
        elif targets.ndim == 4 and targets.shape[1] == 1:
            t3 = targets[:, 0]
        elif targets.ndim == outputs.ndim and targets.shape[1] != 1:
            raise ValueError(f"Target shape {targets.shape} has multiple channels but expected class indices "
                             "[N, H, W] or [N, 1, H, W] for CE.")
        else:
            raise ValueError(f"Unsupported target shape {targets.shape} for CE. Expected [N, H, W] or [N, 1, H, W].")
        return self._run(outputs, t3, self.dice_weight, self.ce_weight, self.smooth_dice, self.ignore_index,
                         self.class_weights)


# ---------------------------------------------------------------------------------------------------------------------
# prompt-model losses (probability inputs)
# ---------------------------------------------------------------------------------------------------------------------
def _classify_nonlin(fn):
    """The reference takes arbitrary callables (``nll_nonlin``, ``dice_nonlin``).  The fused kernel knows two:
    ``None`` (identity) and the prompt notebook's ``lambda x: torch.log(x + eps)`` (prompt_based/prompt.ipynb:68).
    The callable is probed on three values; anything else is refused loudly.  Returns ("identity"|"log", eps)."""
    if fn is None:
        return "identity", 0.0
    probe = torch.tensor([0.0, 0.25, 1.0], dtype=torch.float64)
    try:
        out = fn(probe)
    except Exception as e:
        raise NotImplementedError(f"nonlinearity {fn!r} cannot be evaluated on a CPU probe tensor: {e}")
    if torch.equal(out, probe):
        return "identity", 0.0
    eps = float(torch.exp(out[0]))
    if eps > 0 and torch.allclose(out, torch.log(probe + eps), rtol=1e-9, atol=1e-12):
        return "log", eps
    raise NotImplementedError("only nll_nonlin=None or x -> log(x + eps) is supported by the fused Dice+NLL kernel")


class WeightedMemoryEfficientDiceLossPrompt(_FusedLossBase):
    """Soft Dice over the batch on logits (``apply_softmax=True``) or probabilities (``False``); returns ``-dice``
    (reference: utils/weighted_loss.py:170-273).  ``dice_nonlin`` must be None (the reference's own wrapper never
    forwards it, :299-304)."""

    def __init__(self, dice_nonlin=None, apply_softmax: bool = True, ignore_index: Optional[int] = None,
                 class_weights: Optional[torch.Tensor] = None, smooth: float = 1e-5):
        super().__init__()
        if dice_nonlin is not None and _classify_nonlin(dice_nonlin)[0] != "identity":
            raise NotImplementedError("dice_nonlin is not supported by the fused kernel")
        self.apply_softmax = apply_softmax
        self.ignore_index = ignore_index
        self.smooth = smooth
        self.dice_nonlin = dice_nonlin
        if class_weights is not None:
            assert isinstance(class_weights, torch.Tensor), "class_weights must be a torch.Tensor"
        self.class_weights = class_weights

    def forward(self, x, y):
        if y.dim() == x.dim() - 1 and y.shape == x.shape[:1] + x.shape[2:]:
            y3 = y
        elif y.dim() == x.dim() and y.shape[1] == 1:
            y3 = y[:, 0]
        elif y.dim() == x.dim() and y.shape == x.shape:
            raise NotImplementedError("soft / one-hot float targets are not supported by the fused kernel")
        else:
            raise ValueError(f"Shape mismatch: probs {x.shape}, y {y.shape}")
        kind = L.LOSS_LOGITS if self.apply_softmax else L.LOSS_PROBS_RAW
        return self._run(x, y3, 1.0, 0.0, self.smooth, self.ignore_index, self.class_weights, kind, 0.0)


class WeightedDiceNLLLoss(_FusedLossBase):
    """``dice_weight * SoftDice(p) + nll_weight * NLLLoss(nll_nonlin(p))`` (reference: utils/weighted_loss.py:276-343).

    Supported configurations (anything else raises NotImplementedError):
      * ``apply_softmax=False`` with ``nll_nonlin = lambda x: torch.log(x + eps)`` -- the prompt model's loss
        (prompt_based/prompt.ipynb:68-70): inputs are probabilities;
      * ``apply_softmax=False`` with ``nll_nonlin=None``: NLLLoss on the raw probabilities.
    ``dice_nonlin`` is stored but -- exactly as in the reference (:299-304) -- never applied."""

    def __init__(self, dice_weight: float = 1.0, nll_weight: float = 1.0, ignore_index: Optional[int] = None,
                 class_weights: Optional[torch.Tensor] = None, smooth_dice: float = 1e-5, apply_softmax: bool = True,
                 dice_nonlin=None, nll_nonlin=None, nll_kwargs={}):
        super().__init__()
        self.dice_weight = dice_weight
        self.nll_weight = nll_weight
        self.ignore_index = ignore_index
        self.dice_nonlin = dice_nonlin
        self.nll_nonlin = nll_nonlin
        self.class_weights = class_weights
        self.smooth_dice = smooth_dice
        self.apply_softmax = apply_softmax
        extra = {k: v for k, v in dict(nll_kwargs).items() if not (k == "reduction" and v == "mean")}
        if extra:
            raise NotImplementedError(f"nll_kwargs {sorted(extra)} are not supported by the fused Dice+NLL kernel")
        if apply_softmax:
            raise NotImplementedError("WeightedDiceNLLLoss(apply_softmax=True) feeds NLLLoss with raw logits in the reference; "
                                      "only the probability-input form (apply_softmax=False) is on the accelerated path")
        kind, eps = _classify_nonlin(nll_nonlin)
        self._kind = L.LOSS_PROBS_LOG if kind == "log" else L.LOSS_PROBS_RAW
        self._eps = eps
        self.dice = WeightedMemoryEfficientDiceLossPrompt(apply_softmax=apply_softmax, ignore_index=ignore_index,
                                                          class_weights=class_weights, smooth=smooth_dice)

    def forward(self, outputs, targets):
        if targets.ndim == 3:
            t3 = targets
        elif targets.ndim == 4 and targets.shape[1] == 1:
            t3 = targets[:, 0]
        elif targets.ndim == outputs.ndim and targets.shape[1] != 1:
            raise ValueError(f"Target shape {targets.shape} has multiple channels but expected class indices "
                             "[N, H, W] or [N, 1, H, W] for CE.")
        else:
            raise ValueError(f"Unsupported target shape {targets.shape} for CE. Expected [N, H, W] or [N, 1, H, W].")
        return self._run(outputs, t3, self.dice_weight, self.nll_weight, self.smooth_dice, self.ignore_index,
                         self.class_weights, self._kind, self._eps)
