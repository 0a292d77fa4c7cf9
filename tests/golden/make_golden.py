"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference (run in the
authoring container only: ``python tests/golden/make_golden.py``).

The reference (/root/reference, Python + PyTorch) is imported through ``oracle/ref_shim.py``;
its outputs on seeded synthetic inputs are stored as small ``.npz`` fixtures that travel with the
repo (the GPU box has no /root/reference).  torch version used: see ``meta.json``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from image_segmentation_b200.utils.synthetic import make_batch  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CLASS_W4 = [0.2046795970925636, 1.0271954434416883, 1.2293222812780409, 1.5388026781877073]  # unet/unet.ipynb:57


def tensor_digest(t: torch.Tensor):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()] + t[:4].tolist()
                    + t[-4:].tolist() if t.numel() >= 4 else [t.sum().item()] + t.tolist())


def gen_init(ref):
    out = {}
    for din, dout in ((3, 3), (3, 4), (4, 1)):
        torch.manual_seed(0)
        m = ref.unet(din, dout)
        sd = m.state_dict()
        out[f"keys_{din}{dout}"] = np.array(list(sd.keys()))
        out[f"shapes_{din}{dout}"] = np.array([str(tuple(v.shape)) for v in sd.values()])
        out[f"digest_{din}{dout}"] = np.stack([np.resize(tensor_digest(v), 11) for v in sd.values()])
    np.savez_compressed(os.path.join(OUT, "init.npz"), **out)


def gen_unet_step(ref):
    """One fwd+bwd of unet(3,3) / unet(3,4) / unet(4,1) at 2x32x32 in fp32 and fp64."""
    out = {}
    for din, dout, hw in ((3, 3, 32), (3, 4, 32), (4, 1, 16)):
        tag = f"{din}{dout}"
        x, y = make_batch(2, hw, hw, din, max(dout, 2), seed=1234)
        if dout == 1:
            y = torch.zeros_like(y)
        for dt, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
            torch.manual_seed(0)
            m = ref.unet(din, dout).to(dt).train()
            if dout >= 3:
                w = torch.tensor(CLASS_W4[:dout], dtype=dt)
                loss_fn = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w)
            else:
                loss_fn = ref.WeightedDiceCELoss(smooth_dice=1)
            logits = m(x.to(dt))
            loss = loss_fn(logits, y.squeeze(1))
            loss.backward()
            out[f"logits_{tag}_{dn}"] = logits.detach().numpy()
            out[f"loss_{tag}_{dn}"] = np.array(loss.item())
            names, norms = [], []
            for k, p in m.named_parameters():
                names.append(k)
                norms.append(p.grad.double().norm().item())
            out[f"grad_names_{tag}"] = np.array(names)
            out[f"grad_norms_{tag}_{dn}"] = np.array(norms)
            gd = dict(m.named_parameters())
            for k in ("output.weight", "output.bias", "down1.doubleConvReLU.0.weight", "up4.upsample.bias",
                      "up4.doubleConv.doubleConvReLU.4.weight", "down5.maxpool_doubleConv.1.doubleConvReLU.1.bias"):
                out[f"grad_{tag}_{dn}:{k}"] = gd[k].grad.detach().numpy()
            sd = m.state_dict()
            for k in ("down1.doubleConvReLU.1.running_mean", "down1.doubleConvReLU.1.running_var",
                      "up1.doubleConv.doubleConvReLU.4.running_mean", "up1.doubleConv.doubleConvReLU.4.running_var",
                      "down1.doubleConvReLU.1.num_batches_tracked"):
                out[f"buf_{tag}_{dn}:{k}"] = sd[k].numpy()
            # eval-mode forward after the running-stat update
            m.eval()
            with torch.no_grad():
                out[f"logits_eval_{tag}_{dn}"] = m(x.to(dt)).numpy()
    np.savez_compressed(os.path.join(OUT, "unet_step.npz"), **out)


def gen_loss(ref):
    out = {}
    g = torch.Generator().manual_seed(7)
    cases = []
    for ci, (c, kw) in enumerate([
        (3, dict(smooth_dice=1.0)),
        (3, dict(smooth_dice=1.0, class_weights=CLASS_W4[:3])),
        (4, dict(smooth_dice=1e-5, ignore_index=3)),
        (4, dict(smooth_dice=1.0, class_weights=CLASS_W4, ignore_index=3, dice_weight=0.7, ce_weight=1.3)),
        (4, dict(class_weights=CLASS_W4)),
        (2, dict(smooth_dice=1.0)),
        (1, dict(smooth_dice=1.0)),
    ]):
        logits = (torch.randn(3, c, 12, 10, generator=g) * 2).requires_grad_(True)
        target = torch.randint(0, c, (3, 12, 10), generator=g)
        kw_t = dict(kw)
        if "class_weights" in kw_t:
            kw_t["class_weights"] = torch.tensor(kw_t["class_weights"], dtype=torch.float32)
        fn = ref.WeightedDiceCELoss(**kw_t)
        loss = fn(logits, target)
        loss.backward()
        # same call with the [N,1,H,W] target form (utils/weighted_loss.py:143)
        loss4 = fn(logits.detach(), target.unsqueeze(1))
        assert torch.equal(loss.detach(), loss4)
        out[f"logits_{ci}"] = logits.detach().numpy()
        out[f"target_{ci}"] = target.numpy()
        out[f"loss_{ci}"] = np.array(loss.item())
        out[f"grad_{ci}"] = logits.grad.numpy()
        cases.append(dict(c=c, **kw))
    out["cases"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **out)


def gen_metrics(ref):
    out = {}
    g = torch.Generator().manual_seed(11)
    cases = []
    for ci, (c, ign, shape) in enumerate([(3, None, (17, 23)), (4, 3, (32, 32)), (4, None, (5, 7)), (2, None, (4, 4)),
                                          (3, None, (64, 48))]):
        agg = ref.MetricsHistory(c, ign)
        preds, labels = [], []
        for _ in range(3):
            pred = torch.randn(c, *shape, generator=g)
            # force ties (argmax must pick the lowest class index)
            pred[:, ::3, ::2] = pred[:1, ::3, ::2]
            label = torch.randint(0, c, shape, generator=g)
            agg.accumulate(pred, label)
            preds.append(pred.numpy()); labels.append(label.numpy())
        out[f"pred_{ci}"] = np.stack(preds)
        out[f"label_{ci}"] = np.stack(labels)
        out[f"counts_{ci}"] = np.stack([agg.total_tp.numpy(), agg.total_fp.numpy(), agg.total_fn.numpy(), agg.total_tn.numpy()])
        md, mi, ma = agg.compute_epoch_metrics()
        out[f"means_{ci}"] = np.array([md, mi, ma])
        out[f"perclass_{ci}"] = np.stack([agg.get_last_per_class_dice().numpy(), agg.get_last_per_class_iou().numpy(),
                                          agg.get_last_per_class_acc().numpy()])
        cases.append(dict(c=c, ignore_index=ign))
    out["cases"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)


def gen_curve(ref):
    """100 optimiser steps through the reference's own train_loop (utils/training.py:18-64)."""
    assert ref.training_mod is not None, getattr(ref, "training_error", None)
    out = {}
    for tag, (n, hw, accum) in {"a": (2, 64, 1), "b": (2, 32, 2)}.items():
        torch.manual_seed(0)
        m = ref.unet(3, 3)
        opt = torch.optim.AdamW(m.parameters(), weight_decay=0.01)          # unet/unet.ipynb:80
        w = torch.tensor(CLASS_W4[:3])
        loss_fn = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w)
        batches = [make_batch(n, hw, hw, 3, 3, seed=100 + i, labels="learnable") for i in range(4)]
        steps = 100 if tag == "a" else 20
        losses = []
        import contextlib, io
        for s in range(steps):
            # one "epoch" = `accum` micro-batches = one optimiser step
            loader = [(batches[(s * accum + j) % 4][0], batches[(s * accum + j) % 4][1].to(torch.uint8)) for j in range(accum)]
            with contextlib.redirect_stdout(io.StringIO()):
                avg = ref.training_mod.train_loop(loader, m, loss_fn, opt, accum, torch.device("cpu"), None, hw)
            losses.append(avg)
        out[f"curve_{tag}"] = np.array(losses)
        out[f"cfg_{tag}"] = np.array(json.dumps(dict(n=n, hw=hw, accum=accum, steps=steps)))
    np.savez_compressed(os.path.join(OUT, "curve.npz"), **out)


def gen_eval(ref):
    """Evaluation tail through the reference's own eval_loop / process_batch_reverse (utils/training.py:67-121,
    utils/utils.py:13-115): a stub model returns fixed logits so the fixture pins exactly the crop + bilinear
    resize + per-image loss + confusion counts, not the network."""
    assert ref.training_mod is not None, getattr(ref, "training_error", None)
    import contextlib, io
    tm = ref.training_mod
    out = {}
    g = torch.Generator().manual_seed(21)
    target, c, ign = 48, 4, 3
    sizes = [[(37, 53), (48, 48), (80, 45)], [(31, 97), (64, 40)]]
    w = torch.tensor(CLASS_W4)
    loss_fn = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w, ignore_index=ign)
    batches, all_logits = [], []
    for bi, szs in enumerate(sizes):
        X = [torch.rand(3, h, ww, generator=g) for h, ww in szs]
        y = [torch.randint(0, c, (1, h, ww), generator=g).to(torch.uint8) if i % 2 == 0 else
             torch.randint(0, c, (h, ww), generator=g) for i, (h, ww) in enumerate(szs)]
        logits = torch.randn(len(szs), c, target, target, generator=g) * 2
        logits[:, 1, ::5, ::3] = logits[:, 0, ::5, ::3]          # exact ties survive interpolation where taps coincide
        batches.append((X, y))
        all_logits.append(logits)
        out[f"logits_{bi}"] = logits.numpy()
        for i, lab in enumerate(y):
            out[f"label_{bi}_{i}"] = lab.numpy()
        Xp, metas = tm.process_batch_forward(X, target_size=target)
        out[f"xproc_digest_{bi}"] = tensor_digest(Xp)
        out[f"xproc_{bi}"] = Xp.numpy().astype(np.float32)
        for i, img in enumerate(X):
            out[f"x_{bi}_{i}"] = img.numpy()
        out[f"meta_{bi}"] = np.array(json.dumps(metas))
        for mode in ("bilinear", "nearest"):
            rev = tm.process_batch_reverse(logits, metas, interpolation=mode)
            for i, r in enumerate(rev):
                out[f"rev_{mode}_{bi}_{i}"] = r.numpy()

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.k = 0

        def forward(self, x):
            r = all_logits[self.k]
            self.k += 1
            return r

    agg = ref.MetricsHistory(c, ign)
    with contextlib.redirect_stdout(io.StringIO()):
        avg_loss, mean_dice, mean_iou = tm.eval_loop(batches, Stub(), loss_fn, torch.device("cpu"), target, agg)
    out["result"] = np.array([avg_loss, mean_dice, mean_iou])
    out["counts"] = np.stack([agg.total_tp.numpy(), agg.total_fp.numpy(), agg.total_fn.numpy(), agg.total_tn.numpy()])
    # per-image losses, as the loop computes them
    per = []
    for (X, y), logits in zip(batches, all_logits):
        _, metas = tm.process_batch_forward(X, target_size=target)
        for pred, label in zip(tm.process_batch_reverse(logits, metas, interpolation="bilinear"), y):
            per.append(loss_fn(pred.unsqueeze(0), label.long().unsqueeze(0).squeeze(1)).item())
    out["per_image_loss"] = np.array(per)
    out["cfg"] = np.array(json.dumps(dict(target=target, c=c, ignore_index=ign, sizes=sizes, smooth_dice=1.0)))
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **out)


def start_fixture(seed=31, target=32):
    """Deterministic tiny train / validation loaders shared by gen_start and tests/test_gpu_start.py."""
    g = torch.Generator().manual_seed(seed)
    train = []
    for i in range(3):
        x, y = make_batch(2, target, target, 3, 4, seed=400 + i, labels="learnable")
        train.append((x, y.to(torch.uint8)))
    val = []
    for szs in ([(40, 28), (32, 32)], [(25, 50)]):
        X = [torch.rand(3, h, w, generator=g) for h, w in szs]
        y = [(X[i].mean(0) * 3.999).floor().to(torch.uint8) for i in range(len(szs))]      # labels 0..3 tied to the image
        val.append((X, y))
    return train, val


def gen_start(ref):
    """Two epochs of the reference's start() (utils/training.py:453-617): train_loop + eval_loop + checkpointing."""
    import contextlib, io, tempfile
    tm = ref.training_mod
    train, val = start_fixture()
    torch.manual_seed(0)
    m = ref.unet(3, 4)
    opt = torch.optim.AdamW(m.parameters(), weight_decay=0.01)
    w = torch.tensor(CLASS_W4)
    loss_fn = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w, ignore_index=3)
    agg = ref.MetricsHistory(4, 3)
    buf = io.StringIO()
    # start() pickles the MetricsHistory object: its defining module must be importable under its own name meanwhile
    import types
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.MetricsHistory")}
    pkg = types.ModuleType("utils")
    pkg.MetricsHistory = ref.metrics_mod
    sys.modules["utils"], sys.modules["utils.MetricsHistory"] = pkg, ref.metrics_mod
    try:
        _run_start(tm, ref, m, opt, train, val, loss_fn, agg, buf)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _run_start(tm, ref, m, opt, train, val, loss_fn, agg, buf):
    import contextlib, tempfile
    scratch = os.path.join(ROOT, "build")            # git-ignored scratch directory inside the repository
    os.makedirs(scratch, exist_ok=True)
    with tempfile.TemporaryDirectory(dir=scratch) as d, contextlib.redirect_stdout(buf):
        tm.start(d, "ck.pt", m, opt, train, val, 1, torch.device("cpu"), loss_fn, loss_fn, 32, None, agg, True, True, 4, 3, 2)
        ck = torch.load(os.path.join(d, "ck.pt"), weights_only=True)
        files = sorted(os.listdir(d)) + sorted("metrics/" + f for f in os.listdir(os.path.join(d, "metrics")))
    text = buf.getvalue()
    out = {"stdout": np.array(text), "files": np.array(json.dumps(files)), "ck_keys": np.array(json.dumps(sorted(ck.keys()))),
           "ck_epoch": np.array(ck["epoch"]),
           "ck_best": np.array([ck["best_dev_dice"], ck["best_dev_miou"], ck["best_dev_loss"]]),
           "miou_history": np.array(agg.get_mean_iou_history()), "dice_history": np.array(agg.get_mean_dice_history())}
    np.savez_compressed(os.path.join(OUT, "start.npz"), **out)


def main():
    ref = ref_shim.load()
    torch.set_num_threads(8)
    gen_init(ref)
    gen_unet_step(ref)
    gen_loss(ref)
    gen_metrics(ref)
    gen_curve(ref)
    gen_eval(ref)
    gen_start(ref)
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, threads=torch.get_num_threads(),
                       reference="in5omnia/Image_Segmentation @ /root/reference (unmodified, CPU)"), f, indent=1)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
