"""Synthetic inputs of the shape the reference trains on (no dataset exists in the container).

Images are U[0,1) like ``utils/dataset.py:39`` (jpeg / 255); labels are trimap-like class maps.
Two label recipes:

* ``iid``        i.i.d. uniform classes (BASELINE.md section 3 / SURVEY.md section 8(d) config 1)
* ``learnable``  quantile-thresholded 9x9 box blur of the channel mean, so that a few optimiser
                 steps visibly reduce the loss (used for loss-curve parity)

Everything is generated on the CPU with an explicit ``torch.Generator`` so that the oracle, the
golden-vector script and the CUDA path see bit-identical inputs.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def make_batch(n: int, h: int, w: int, din: int = 3, num_classes: int = 3, seed: int = 1234,
               labels: str = "iid"):
    """Returns (X [n,din,h,w] float32 in [0,1), y [n,1,h,w] int64 in [0,num_classes))."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, din, h, w, generator=g)
    if labels == "iid":
        y = torch.randint(0, num_classes, (n, 1, h, w), generator=g)
    elif labels == "learnable":
        m = F.avg_pool2d(x.mean(1, keepdim=True), 9, stride=1, padding=4, count_include_pad=False)
        qs = torch.linspace(0, 1, num_classes + 1)[1:-1]
        thr = torch.quantile(m.flatten(), qs)
        y = torch.bucketize(m, thr).long()
    else:
        raise ValueError(labels)
    return x, y
