"""Drop-in for the reference's ``utils/MetricsHistory.py``: per-class TP/FP/FN/TN accumulation and
Dice / IoU / accuracy epoch metrics, same public surface (``accumulate``, ``reset``,
``compute_epoch_metrics``, ``to``, ``get_*``, picklable).

Difference in *how*: ``accumulate`` (reference :55-86 = argmax + two one_hot + four masked sums + four
``.cpu()`` syncs per image) is one CUDA kernel that adds exact int64 counts to device-resident
accumulators; nothing is copied to the host until the totals are read (``total_tp`` ... properties,
``compute_epoch_metrics`` or pickling).  Counts are integers, so results are bit-identical to the
reference's float64 sums.
"""
import torch

from .. import _lib as L

_MAX_CLASSES = 8


class MetricsHistory:
    """
    Accumulates TP, FP, FN, TN over an epoch for multi-class segmentation
    and computes Dice, IoU, and Accuracy metrics.
    """

    _HISTORIES = ("mean_dice", "mean_iou", "mean_acc", "per_class_dice", "per_class_iou", "per_class_acc")

    def __init__(self, num_classes: int, ignore_index: int = None, device: str = 'cpu'):
        self.num_classes, self.ignore_index = num_classes, ignore_index
        self._host = torch.zeros(4, num_classes, dtype=torch.float64)      # synced totals: rows tp, fp, fn, tn
        self._pending = {}                                                  # device -> (int64 [4,C], status)
        for name in self._HISTORIES:                                        # epoch_mean_dice_history, ... (reference :26-38)
            setattr(self, f"epoch_{name}_history", [])
        for kind in ("iou", "dice", "acc"):
            setattr(self, f"last_per_class_{kind}", None)
        keep = torch.ones(num_classes, dtype=torch.bool)
        if ignore_index is not None and 0 <= ignore_index < num_classes:
            keep[ignore_index] = False                                      # the class left out of the macro averages
        self.mask = keep

    # ---- device-resident accumulation ----------------------------------------------------------
    def _sync(self):
        """Fold device counters into the float64 host totals (one D2H copy per device)."""
        for dev, (counts, status) in list(self._pending.items()):
            both = torch.cat([counts.flatten(), status.to(torch.int64)]).cpu()
            if int(both[-1]) != 0:
                counts.zero_()
                status.zero_()
                raise RuntimeError("Class values must be smaller than num_classes.")
            self._host += both[:-1].view(4, self.num_classes).to(torch.float64)
            counts.zero_()

    def _totals(self, i):
        self._sync()
        return self._host[i]

    def _set_totals(self, i, value):
        """``agg.total_tp = tensor`` (plain attributes in the reference, :21-24): device counters are folded in first."""
        self._sync()
        self._host[i].copy_(torch.as_tensor(value, dtype=torch.float64).reshape(self.num_classes).cpu())

    total_tp = property(lambda self: self._totals(0), lambda self, v: self._set_totals(0, v))
    total_fp = property(lambda self: self._totals(1), lambda self, v: self._set_totals(1, v))
    total_fn = property(lambda self: self._totals(2), lambda self, v: self._set_totals(2, v))
    total_tn = property(lambda self: self._totals(3), lambda self, v: self._set_totals(3, v))

    def reset(self):
        """Resets the accumulated TP, FP, FN, TN counts."""
        for counts, status in self._pending.values():
            counts.zero_()
        self._host.zero_()

    def _device_counters(self, dev):
        """(int64 [4,C] counts, int32 [1] bad-label flag) resident on ``dev``; kernels add into them."""
        if dev not in self._pending:
            self._pending[dev] = (torch.zeros(4, self.num_classes, dtype=torch.int64, device=dev),
                                  torch.zeros(1, dtype=torch.int32, device=dev))
        return self._pending[dev]

    def accumulate(self, pred: torch.Tensor, label: torch.Tensor):
        """pred: logits or probabilities (C,H,W) [or (N,C,H,W) for a whole batch]; label: (H,W), (1,H,W) [or (N,H,W)]."""
        L.require_cuda(pred, label)
        c = self.num_classes
        if c > _MAX_CLASSES:
            raise NotImplementedError(f"up to {_MAX_CLASSES} classes supported")
        if pred.dim() == 3:
            pred = pred.unsqueeze(0)
        if pred.dim() != 4 or pred.shape[1] != c:
            raise RuntimeError(f"pred must be [C,H,W] with C={c}, got {tuple(pred.shape)}")
        n, _, h, w = pred.shape
        if label.numel() != n * h * w:
            raise RuntimeError(f"label shape {tuple(label.shape)} does not match pred {tuple(pred.shape)}")
        pred = pred if (pred.dtype == torch.float32 and pred.is_contiguous()) else pred.float().contiguous()
        label = label if (label.dtype == torch.int64 and label.is_contiguous()) else label.long().contiguous()
        dev = pred.device
        counts, status = self._device_counters(dev)
        with torch.cuda.device(dev):
            L.argmax_confusion(pred, label, n, c, h, w, counts, None, status)

    def compute_epoch_metrics(self, epsilon: float = 1e-6):
        """Macro-averaged (mean_dice, mean_iou, mean_acc) of the accumulated epoch; appended to the histories.
        (`epsilon` is accepted and unused, as in the reference: an absent class yields NaN.)"""
        self._sync()
        tp, fp, fn, tn = self._host
        union = tp + fp + fn
        per_class = {"iou": tp / union, "dice": (2 * tp) / (tp + union), "acc": (tp + tn) / (union + tn)}
        means = {}
        for kind, values in per_class.items():
            means[kind] = values[self.mask.cpu()].mean().item()
            getattr(self, f"epoch_mean_{kind}_history").append(means[kind])
            getattr(self, f"epoch_per_class_{kind}_history").append(values.numpy())
            setattr(self, f"last_per_class_{kind}", values)
        return means["dice"], means["iou"], means["acc"]

    def to(self, device):
        """Moves what the reference moves that a caller can observe (:130-150): the class mask and the last per-class
        metric tensors.  The count totals stay float64 host tensors fed by exact device counters that follow the
        predictions' device (no per-image copies, unlike the reference's ``.cpu()`` x 4 per image, :83-86)."""
        self.mask = self.mask.to(device)
        for kind in ("iou", "dice", "acc"):
            v = getattr(self, f"last_per_class_{kind}")
            if v is not None:
                setattr(self, f"last_per_class_{kind}", v.to(device))

    def all_reduce(self, group=None):
        """Data-parallel evaluation: sum the exact counts over ranks (one small collective)."""
        import torch.distributed as dist
        self._sync()
        if dist.is_available() and dist.is_initialized():
            t = self._host.clone()
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.all_reduce(t, group=group)
            self._host.copy_(t.cpu())

    def __getstate__(self):
        self._sync()
        state = self.__dict__.copy()
        state["_pending"] = {}
        return state

    def __setstate__(self, state):
        # torch.load(..., map_location=device) moves every pickled tensor (utils/training.py:351); the float64 totals
        # are host tensors by construction
        self.__dict__.update(state)
        self._host = self._host.cpu()
        self._pending = {}

    def get_ignore_index(self):
        return self.ignore_index

    def get_num_classes(self):
        return self.num_classes


def _install_getters():
    """get_mean_*_history / get_class_*_history / get_last_per_class_* of the reference (:152-182), generated."""
    for kind in ("dice", "iou", "acc"):
        for public, attr in ((f"get_mean_{kind}_history", f"epoch_mean_{kind}_history"),
                             (f"get_class_{kind}_history", f"epoch_per_class_{kind}_history"),
                             (f"get_last_per_class_{kind}", f"last_per_class_{kind}")):
            setattr(MetricsHistory, public, (lambda a: lambda self: getattr(self, a))(attr))


_install_getters()
