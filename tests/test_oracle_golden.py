"""Pins the CPU oracle (oracle/) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py) and, when /root/reference is present, against the live
reference.  CPU only."""
import json

import numpy as np
import pytest
import torch

from oracle import eval_oracle, loss_oracle, metrics_oracle, ref_shim, unet_oracle
from image_segmentation_b200.utils.synthetic import make_batch

CLASS_W4 = [0.2046795970925636, 1.0271954434416883, 1.2293222812780409, 1.5388026781877073]


def _digest(t):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()] + t[:4].tolist() + t[-4:].tolist()
                    if t.numel() >= 4 else [t.sum().item()] + t.tolist())


@pytest.mark.parametrize("din,dout", [(3, 3), (3, 4), (4, 1)])
def test_init_matches_reference_rng_stream(golden, din, dout):
    g = golden["init"]
    torch.manual_seed(0)
    sd = unet_oracle.init_state_dict(din, dout)
    tag = f"{din}{dout}"
    assert list(sd.keys()) == list(g[f"keys_{tag}"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g[f"shapes_{tag}"])
    assert len(sd) == 136
    assert sum(v.numel() for k, v in sd.items() if k in unet_oracle.param_names(sd)) == {
        (3, 3): 31043651, (3, 4): 31043716, (4, 1): 31044097}[(din, dout)]
    dig = np.stack([np.resize(_digest(v), 11) for v in sd.values()])
    np.testing.assert_array_equal(dig, g[f"digest_{tag}"])      # bit-exact: same RNG stream, same bounds


def _loss_fn(dout, dtype):
    kw = dict(smooth_dice=1.0)
    if dout >= 3:
        kw["class_weights"] = torch.tensor(CLASS_W4[:dout], dtype=dtype)
    return lambda logits, y: loss_oracle.dice_ce_loss(logits, y, dtype=dtype, **kw)


@pytest.mark.parametrize("din,dout,hw", [(3, 3, 32), (3, 4, 32), (4, 1, 16)])
def test_unet_step_matches_golden(golden, din, dout, hw):
    g = golden["unet_step"]
    tag = f"{din}{dout}"
    x, y = make_batch(2, hw, hw, din, max(dout, 2), seed=1234)
    if dout == 1:
        y = torch.zeros_like(y)
    for dt, dn, tol in ((torch.float64, "f64", 1e-9), (torch.float32, "f32", 2e-4)):
        torch.manual_seed(0)
        sd = unet_oracle.init_state_dict(din, dout)
        loss, logits, grads, newbuf = unet_oracle.loss_and_grads(sd, x, y.squeeze(1), _loss_fn(dout, dt), dtype=dt)
        ref_logits = g[f"logits_{tag}_{dn}"]
        err = np.abs(logits.numpy() - ref_logits).max() / np.abs(ref_logits).max()
        assert err < tol, (dn, err)
        assert abs(loss.item() - float(g[f"loss_{tag}_{dn}"])) < tol * max(1.0, abs(loss.item()))
        if dn == "f64":
            # gradients: only the fp64 run is compared tightly (fp32 carries ReLU/maxpool flips)
            names = list(g[f"grad_names_{tag}"])
            norms = g[f"grad_norms_{tag}_{dn}"]
            for k, n_ref in zip(names, norms):
                n = grads[k].double().norm().item()
                if k.endswith((".0.bias", ".3.bias")) and "doubleConvReLU" in k:
                    continue  # conv bias before train-mode BN: gradient is rounding noise
                assert abs(n - n_ref) <= 1e-7 * max(n_ref, 1e-12) + 1e-12, (k, n, n_ref)
            for key in g.files:
                if key.startswith(f"grad_{tag}_{dn}:"):
                    k = key.split(":", 1)[1]
                    np.testing.assert_allclose(grads[k].numpy(), g[key], rtol=1e-6, atol=1e-10)
            for key in g.files:
                if key.startswith(f"buf_{tag}_{dn}:"):
                    k = key.split(":", 1)[1]
                    np.testing.assert_allclose(newbuf[k].numpy(), g[key], rtol=1e-9, atol=1e-12)
            # eval-mode forward with the updated running statistics
            sd2 = {k: v.double() if v.is_floating_point() else v for k, v in sd.items()}
            sd2.update(newbuf)
            ev = unet_oracle.forward(sd2, x.double(), training=False)
            np.testing.assert_allclose(ev.numpy(), g[f"logits_eval_{tag}_{dn}"], rtol=1e-8, atol=1e-10)


def test_loss_matches_golden(golden):
    g = golden["loss"]
    cases = json.loads(str(g["cases"]))
    for ci, case in enumerate(cases):
        kw = {k: v for k, v in case.items() if k != "c"}
        if "class_weights" in kw:
            kw["class_weights"] = torch.tensor(kw["class_weights"], dtype=torch.float32)
        logits = torch.from_numpy(g[f"logits_{ci}"])
        target = torch.from_numpy(g[f"target_{ci}"])
        loss = loss_oracle.dice_ce_loss(logits, target, **kw)
        assert abs(loss.item() - float(g[f"loss_{ci}"])) < 2e-6, (ci, loss.item(), float(g[f"loss_{ci}"]))
        grad = loss_oracle.dice_ce_grad(logits, target, **kw)
        np.testing.assert_allclose(grad.numpy(), g[f"grad_{ci}"], rtol=2e-4, atol=2e-8)
        # analytic gradient == autograd of the closed form
        lg = logits.double().requires_grad_(True)
        loss_oracle.dice_ce_loss(lg, target, **kw).backward()
        np.testing.assert_allclose(grad.numpy(), lg.grad.numpy(), rtol=1e-9, atol=1e-14)
        # [N,1,H,W] targets are accepted, other shapes raise (utils/weighted_loss.py:141-149)
        assert loss_oracle.dice_ce_loss(logits, target.unsqueeze(1), **kw).item() == loss.item()
        with pytest.raises(ValueError):
            loss_oracle.dice_ce_loss(logits, target.unsqueeze(1).repeat(1, 2, 1, 1), **kw)


def test_metrics_match_golden(golden):
    g = golden["metrics"]
    cases = json.loads(str(g["cases"]))
    for ci, case in enumerate(cases):
        c, ign = case["c"], case["ignore_index"]
        tot = np.zeros((4, c), dtype=np.int64)
        for pred, label in zip(g[f"pred_{ci}"], g[f"label_{ci}"]):
            tot += np.stack(metrics_oracle.confusion_counts(pred, label, c))
        np.testing.assert_array_equal(tot.astype(np.float64), g[f"counts_{ci}"])   # integer-exact
        md, mi, ma, pd_, pi, pa = metrics_oracle.epoch_metrics(*tot, ignore_index=ign)
        np.testing.assert_allclose([md, mi, ma], g[f"means_{ci}"], rtol=1e-12)
        np.testing.assert_allclose(np.stack([pd_, pi, pa]), g[f"perclass_{ci}"], rtol=1e-12)
    with pytest.raises(RuntimeError):
        metrics_oracle.confusion_counts(np.zeros((3, 2, 2), np.float32), np.full((2, 2), 3), 3)


def test_argmax_nan_and_ties():
    pred = np.array([[[1.0, np.nan, 2.0]], [[1.0, 5.0, np.nan]], [[0.5, np.nan, 2.0]]], dtype=np.float32)
    ref = torch.argmax(torch.from_numpy(pred), dim=0).numpy()
    np.testing.assert_array_equal(metrics_oracle.argmax_first(pred), ref)


def test_curve_oracle_module_reproduces_reference_train_loop(golden):
    """First 5 optimiser steps of the golden 100-step curve, re-run with the oracle module."""
    g = golden["curve"]
    cfg = json.loads(str(g["cfg_b"]))
    n, hw, accum = cfg["n"], cfg["hw"], cfg["accum"]
    torch.manual_seed(0)
    m = unet_oracle.OracleUNet(3, 3).train()
    opt = torch.optim.AdamW(m.parameters(), weight_decay=0.01)
    w = torch.tensor(CLASS_W4[:3])
    batches = [make_batch(n, hw, hw, 3, 3, seed=100 + i, labels="learnable") for i in range(4)]
    losses = []
    for s in range(5):
        opt.zero_grad()
        for j in range(accum):
            x, y = batches[(s * accum + j) % 4]
            loss = loss_oracle.dice_ce_loss(m(x), y, smooth_dice=1.0, class_weights=w, dtype=torch.float32)
            (loss / accum).backward()
        opt.step()
        losses.append(loss.item())
    np.testing.assert_allclose(losses, g["curve_b"][:5], rtol=2e-3)
    assert g["curve_a"][-1] < g["curve_a"][0] - 0.3      # the learnable recipe really learns


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
def test_oracle_against_live_reference():
    ref = ref_shim.load()
    x, y = make_batch(2, 32, 32, 3, 4, seed=5)
    torch.manual_seed(3)
    m = ref.unet(3, 4).double().train()
    torch.manual_seed(3)
    sd = unet_oracle.init_state_dict(3, 4)
    for k, v in m.state_dict().items():
        assert torch.equal(v.float() if v.is_floating_point() else v, sd[k]), k
    w = torch.tensor(CLASS_W4, dtype=torch.float64)
    fn = ref.WeightedDiceCELoss(smooth_dice=1e-5, class_weights=w, ignore_index=3)
    logits_ref = m(x.double())
    loss_ref = fn(logits_ref, y.squeeze(1))
    loss_ref.backward()
    loss, logits, grads, _ = unet_oracle.loss_and_grads(
        sd, x, y.squeeze(1),
        lambda lg, t: loss_oracle.dice_ce_loss(lg, t, smooth_dice=1e-5, class_weights=w, ignore_index=3),
        dtype=torch.float64)
    np.testing.assert_allclose(logits.numpy(), logits_ref.detach().numpy(), rtol=1e-9, atol=1e-11)
    assert abs(loss.item() - loss_ref.item()) < 1e-10
    for k, p in m.named_parameters():
        if "doubleConvReLU" in k and k.endswith((".0.bias", ".3.bias")):
            continue
        np.testing.assert_allclose(grads[k].numpy(), p.grad.numpy(), rtol=1e-6, atol=1e-11, err_msg=k)


def test_native_ops_variant_equals_closed_form():
    """The timing variant of the oracle (F.batch_norm / F.conv_transpose2d, as the reference calls them) computes the same thing."""
    x, y = make_batch(2, 32, 32, 3, 3, seed=8)
    torch.manual_seed(1)
    a = unet_oracle.OracleUNet(3, 3).double().train()
    torch.manual_seed(1)
    b = unet_oracle.OracleUNet(3, 3, native_ops=True).double().train()
    ya, yb = a(x.double()), b(x.double())
    np.testing.assert_allclose(ya.detach().numpy(), yb.detach().numpy(), rtol=1e-9, atol=1e-11)
    for (ka, va), (kb, vb) in zip(a.reference_state_dict().items(), b.reference_state_dict().items()):
        np.testing.assert_allclose(va.double().numpy(), vb.double().numpy(), rtol=1e-9, atol=1e-12, err_msg=ka)


def test_eval_tail_oracle_matches_reference_golden(golden):
    """oracle/eval_oracle.py against the reference's process_batch_forward/reverse + eval_loop (gen_eval)."""
    g = golden["eval"]
    cfg = json.loads(str(g["cfg"]))
    kw = dict(smooth_dice=cfg["smooth_dice"], class_weights=torch.tensor(CLASS_W4), ignore_index=cfg["ignore_index"])
    batches = []
    for bi, szs in enumerate(cfg["sizes"]):
        metas = json.loads(str(g[f"meta_{bi}"]))
        for i, (h, w) in enumerate(szs):
            m = eval_oracle.resize_meta(h, w, cfg["target"])
            assert tuple(m["new_size"]) == tuple(metas[i]["new_size"]) and tuple(m["pad"]) == tuple(metas[i]["pad"])
            assert m["scale"] == metas[i]["scale"]
            for mode, tol in (("bilinear", 2e-6), ("nearest", 0.0)):
                r = eval_oracle.crop_resize(g[f"logits_{bi}"][i], metas[i], mode)
                np.testing.assert_allclose(r, g[f"rev_{mode}_{bi}_{i}"], rtol=0, atol=tol)
        batches.append((g[f"logits_{bi}"], metas, [g[f"label_{bi}_{i}"] for i in range(len(szs))]))
    avg, md, mi, counts = eval_oracle.eval_epoch(batches, cfg["c"], kw, cfg["ignore_index"])
    np.testing.assert_array_equal(counts.astype(np.float64), g["counts"])
    np.testing.assert_allclose([avg, md, mi], g["result"], rtol=1e-6)
    per = np.concatenate([eval_oracle.eval_batch(*b, cfg["c"], kw)[0] for b in batches])
    np.testing.assert_allclose(per, g["per_image_loss"], rtol=2e-6)


def test_process_batch_forward_matches_reference_golden(golden):
    """Host-side input preparation (resize with antialiasing + centred zero padding, utils/utils.py:13-49,77-99)."""
    from image_segmentation_b200.utils.utils import process_batch_forward
    g = golden["eval"]
    cfg = json.loads(str(g["cfg"]))
    for bi, szs in enumerate(cfg["sizes"]):
        X = [torch.from_numpy(g[f"x_{bi}_{i}"]) for i in range(len(szs))]
        Xp, metas = process_batch_forward(X, target_size=cfg["target"])
        np.testing.assert_allclose(Xp.numpy(), g[f"xproc_{bi}"], rtol=0, atol=1e-6)
        want = json.loads(str(g[f"meta_{bi}"]))
        for m, w in zip(metas, want):
            assert tuple(m["pad"]) == tuple(w["pad"]) and tuple(m["new_size"]) == tuple(w["new_size"])
            assert tuple(m["original_size"]) == tuple(w["original_size"]) and m["scale"] == w["scale"]
