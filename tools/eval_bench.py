"""Evaluation-tail micro benchmark: fused ragged-batch kernel vs the per-image sequence the
reference runs (crop + F.interpolate + loss + metrics + .item() per image), both on the GPU.

    python tools/eval_bench.py [--images 64] [--target 256] [--classes 4]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_segmentation_b200.utils.MetricsHistory import MetricsHistory  # noqa: E402
from image_segmentation_b200.utils.training import _EvalTail  # noqa: E402
from image_segmentation_b200.utils.utils import process_batch_reverse  # noqa: E402
from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss  # noqa: E402


def meta_for(h, w, t):
    scale = min(t / w, t / h)
    nw, nh = int(round(w * scale)), int(round(h * scale))
    pw, ph = t - nw, t - nh
    return {"original_size": (h, w), "new_size": (nh, nw), "pad": (pw // 2, ph // 2, pw - pw // 2, ph - ph // 2),
            "scale": scale}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--target", type=int, default=256)
    ap.add_argument("--classes", type=int, default=4)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    # Oxford-IIIT-Pet-like sizes (the reference's dataset): around 500x375, both orientations
    sizes = [(int(torch.randint(300, 520, (1,), generator=g)), int(torch.randint(300, 520, (1,), generator=g)))
             for _ in range(a.images)]
    metas = [meta_for(h, w, a.target) for h, w in sizes]
    logits = torch.randn(a.images, a.classes, a.target, a.target, generator=g).to(dev)
    labels = [torch.randint(0, a.classes, (h, w), generator=g).to(torch.uint8) for h, w in sizes]
    labels_dev = [l.to(dev) for l in labels]
    cw = torch.rand(a.classes, generator=g) + 0.5
    loss_fn = WeightedDiceCELoss(smooth_dice=1.0, class_weights=cw, ignore_index=a.classes - 1)
    out_pixels = sum(h * w for h, w in sizes)

    def fused():
        agg = MetricsHistory(a.classes, a.classes - 1)
        tail = _EvalTail(loss_fn, agg, dev)
        tail.batch(logits, metas, labels_dev)
        return tail.finish()

    def per_image():
        # what utils/training.py:91-101 does with this package's per-op kernels (one loss + one metrics call and one
        # .item() per image)
        agg = MetricsHistory(a.classes, a.classes - 1)
        tot = 0.0
        for pred, lab in zip(process_batch_reverse(logits, metas), labels_dev):
            tot += loss_fn(pred.unsqueeze(0), lab.long().unsqueeze(0)).item()
            agg.accumulate(pred, lab.long())
        return tot

    def torch_ref():
        # the reference's own op sequence on the GPU with stock torch kernels
        ce = torch.nn.CrossEntropyLoss(weight=cw.to(dev), ignore_index=a.classes - 1)
        tot = 0.0
        for i, (m, lab) in enumerate(zip(metas, labels_dev)):
            l, t, _, _ = m["pad"]
            nh, nw = m["new_size"]
            pred = F.interpolate(logits[i:i + 1, :, t:t + nh, l:l + nw], size=m["original_size"], mode="bilinear",
                                 align_corners=False)
            y = lab.long().unsqueeze(0)
            p = torch.softmax(pred, 1)
            oh = torch.zeros_like(p).scatter_(1, y.unsqueeze(1), 1)
            inter, sp, sg = (p * oh).sum((0, 2, 3)), p.sum((0, 2, 3)), oh.sum((0, 2, 3))
            dc = (2 * inter + 1.0) / torch.clip(sp + sg + 1.0, 1e-8)
            tot += (-(dc[:-1] * cw.to(dev)[:-1]).sum() / cw[:-1].sum() + ce(pred, y)).item()
            hard = pred[0].argmax(0)
            ph, lh = F.one_hot(hard, a.classes).bool(), F.one_hot(y[0], a.classes).bool()
            _ = [(ph & lh).sum((0, 1)).cpu(), (ph & ~lh).sum((0, 1)).cpu(), (~ph & lh).sum((0, 1)).cpu(),
                 (~ph & ~lh).sum((0, 1)).cpu()]
        return tot

    res = {}
    for name, fn in (("fused", fused), ("per_image_unetk", per_image), ("per_image_torch", torch_ref)):
        v = fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.reps if name == "fused" else max(2, a.reps // 5)):
            fn()
        torch.cuda.synchronize()
        n = a.reps if name == "fused" else max(2, a.reps // 5)
        res[name] = {"ms_per_batch": (time.perf_counter() - t0) / n * 1e3, "loss_sum": v}
    # device time of the fused launch alone
    agg = MetricsHistory(a.classes, a.classes - 1)
    tail = _EvalTail(loss_fn, agg, dev)
    packed = torch.cat([l.reshape(-1) for l in labels_dev])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tail.batch(logits, metas, labels_dev)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        tail.batch(logits, metas, labels_dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    bytes_alg = out_pixels * 1 + logits.numel() * 4
    res["fused_device_ms_incl_host_prep"] = ms
    res["algorithmic_GBps"] = bytes_alg / ms / 1e6
    res["images"], res["out_pixels"], res["packed_label_bytes"] = a.images, out_pixels, packed.numel()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
