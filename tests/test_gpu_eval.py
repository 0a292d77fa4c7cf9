"""Parity of the evaluation tail: crop + bilinear resize of the logits back to each image's
original size, per-image Dice+CE loss, confusion counts -- CUDA path (through the C ABI) against the CPU oracle
(bit-exact for the interpolation and the integer counts) and against golden vectors produced by the unmodified
reference's process_batch_reverse / eval_loop (tests/golden/make_golden.py:gen_eval).

Tolerances: interpolation bit-exact vs the oracle and 2e-6 abs vs ATen's CPU kernel (which contracts FMAs);
loss 1e-5 relative (fp32 softmax, float/double reductions in a different order); counts exact.
"""
import contextlib
import io
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from image_segmentation_b200.utils.MetricsHistory import MetricsHistory  # noqa: E402
from image_segmentation_b200.utils.training import eval_loop  # noqa: E402
from image_segmentation_b200.utils.utils import process_batch_reverse, reverse_resize_and_padding  # noqa: E402
from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss  # noqa: E402
from oracle import eval_oracle  # noqa: E402

DEV = "cuda"
CLASS_W4 = [0.2046795970925636, 1.0271954434416883, 1.2293222812780409, 1.5388026781877073]


class Stub(torch.nn.Module):
    """Returns prepared logits batch after batch, so the tests pin the tail, not the network."""

    def __init__(self, batches):
        super().__init__()
        self.batches, self.k = batches, 0

    def forward(self, x):
        r = self.batches[self.k].to(x.device)
        self.k += 1
        return r


def _golden_batches(g):
    cfg = json.loads(str(g["cfg"]))
    out = []
    for bi, szs in enumerate(cfg["sizes"]):
        metas = json.loads(str(g[f"meta_{bi}"]))
        for m in metas:
            for k in ("original_size", "new_size", "pad"):
                m[k] = tuple(m[k])
        labels = [torch.from_numpy(g[f"label_{bi}_{i}"]) for i in range(len(szs))]
        images = [torch.from_numpy(g[f"x_{bi}_{i}"]) for i in range(len(szs))]
        out.append((torch.from_numpy(g[f"logits_{bi}"]), metas, labels, images))
    return cfg, out


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
def test_process_batch_reverse_matches_oracle_bitwise_and_reference(golden, mode):
    g = golden["eval"]
    _, batches = _golden_batches(g)
    for bi, (logits, metas, _, _) in enumerate(batches):
        got = process_batch_reverse(logits.to(DEV), metas, interpolation=mode)
        for i, (r, m) in enumerate(zip(got, metas)):
            want = eval_oracle.crop_resize(logits[i].numpy(), m, mode)
            assert tuple(r.shape) == want.shape
            np.testing.assert_array_equal(r.cpu().numpy(), want)                           # bit-exact
            np.testing.assert_allclose(r.cpu().numpy(), g[f"rev_{mode}_{bi}_{i}"], rtol=0, atol=2e-6)
        single = reverse_resize_and_padding(logits[0].to(DEV), metas[0], mode)
        assert torch.equal(single, got[0])


def test_eval_loop_matches_reference_golden(golden):
    g = golden["eval"]
    cfg, batches = _golden_batches(g)
    loader = [(images, labels) for _, _, labels, images in batches]
    model = Stub([b[0] for b in batches])
    loss_fn = WeightedDiceCELoss(smooth_dice=cfg["smooth_dice"], class_weights=torch.tensor(CLASS_W4),
                                 ignore_index=cfg["ignore_index"])
    agg = MetricsHistory(cfg["c"], cfg["ignore_index"])
    with contextlib.redirect_stdout(io.StringIO()) as out:
        avg_loss, mean_dice, mean_iou = eval_loop(loader, model, loss_fn, torch.device(DEV), cfg["target"], agg)
    want = g["result"]
    assert abs(avg_loss - want[0]) < 1e-5 * abs(want[0])
    np.testing.assert_array_equal(torch.stack([agg.total_tp, agg.total_fp, agg.total_fn, agg.total_tn]).numpy(),
                                  g["counts"])
    np.testing.assert_allclose([mean_dice, mean_iou], want[1:], rtol=1e-12)
    assert "Images Processed: 5" in out.getvalue() and "Class 3:" in out.getvalue()


def test_fused_tail_equals_generic_tail_and_per_image_losses(golden):
    """The fused kernel, the generic path (resized predictions -> loss / metrics objects per image) and the golden
    per-image losses of the reference agree."""
    g = golden["eval"]
    cfg, batches = _golden_batches(g)
    from image_segmentation_b200.utils.training import _EvalTail
    loss_fn = WeightedDiceCELoss(smooth_dice=1.0, class_weights=torch.tensor(CLASS_W4), ignore_index=3)
    agg_f, agg_g = MetricsHistory(4, 3), MetricsHistory(4, 3)
    tail = _EvalTail(loss_fn, agg_f, DEV)
    generic = []
    for logits, metas, labels, _ in batches:
        lg = logits.to(DEV)
        tail.batch(lg, metas, labels)
        for pred, label in zip(process_batch_reverse(lg, metas), labels):
            label = label.to(DEV).long()
            generic.append(loss_fn(pred.unsqueeze(0), label.reshape(1, *pred.shape[1:])).item())
            agg_g.accumulate(pred, label)
    fused = torch.cat(tail.per_image).cpu().numpy()
    np.testing.assert_allclose(fused, g["per_image_loss"], rtol=1e-5)
    np.testing.assert_allclose(generic, g["per_image_loss"], rtol=1e-5)
    assert abs(tail.finish() - float(np.sum(fused.astype(np.float64)))) < 1e-12
    assert torch.equal(agg_f.total_tp, agg_g.total_tp) and torch.equal(agg_f.total_tn, agg_g.total_tn)
    assert torch.equal(agg_f.total_fp, agg_g.total_fp) and torch.equal(agg_f.total_fn, agg_g.total_fn)


@pytest.mark.parametrize("c,label_dtype,weights,ignore", [(3, torch.uint8, False, None), (4, torch.int64, True, 3),
                                                          (1, torch.int64, False, None), (2, torch.uint8, True, 0)])
def test_ragged_batch_against_oracle(c, label_dtype, weights, ignore):
    gen = torch.Generator().manual_seed(100 + c)
    target = 256
    sizes = [(300, 200), (256, 256), (97, 401), (500, 333), (64, 64), (1, 7), (255, 257)]
    metas = [eval_oracle.resize_meta(h, w, target) for h, w in sizes]
    logits = torch.randn(len(sizes), c, target, target, generator=gen) * 3
    hi = max(c, 2) if c > 1 else 2
    labels = [torch.randint(0, hi if c > 1 else 1, (h, w), generator=gen).to(label_dtype) for h, w in sizes]
    cw = torch.tensor(CLASS_W4[:c]) if weights else None
    kw = dict(smooth_dice=1e-5, class_weights=cw, ignore_index=ignore)
    want_losses, want_counts, _ = eval_oracle.eval_batch(logits.numpy(), metas, [l.numpy() for l in labels], c, kw)

    from image_segmentation_b200.utils.training import _EvalTail
    loss_fn = WeightedDiceCELoss(smooth_dice=1e-5, class_weights=cw, ignore_index=ignore)
    agg = MetricsHistory(c, ignore)
    tail = _EvalTail(loss_fn, agg, DEV)
    tail.batch(logits.to(DEV), metas, labels)
    got = torch.cat(tail.per_image).cpu().numpy()
    np.testing.assert_allclose(got, want_losses, rtol=2e-5, atol=1e-6)
    np.testing.assert_array_equal(torch.stack([agg.total_tp, agg.total_fp, agg.total_fn, agg.total_tn]).numpy()
                                  .astype(np.int64), want_counts)
    total = sum(h * w for h, w in sizes)
    assert int((agg.total_tp + agg.total_fn)[..., :].sum()) == total     # every pixel has exactly one true class


def test_bad_label_raises_like_the_reference():
    metas = [eval_oracle.resize_meta(20, 30, 32)]
    loss_fn = WeightedDiceCELoss(smooth_dice=1.0)
    agg = MetricsHistory(3)
    from image_segmentation_b200.utils.training import _EvalTail
    tail = _EvalTail(loss_fn, agg, DEV)
    lab = torch.zeros(20, 30, dtype=torch.uint8)
    lab[3, 4] = 255
    tail.batch(torch.randn(1, 3, 32, 32, device=DEV), metas, [lab])
    with pytest.raises(RuntimeError):
        agg.compute_epoch_metrics()
    with pytest.raises(ValueError):
        tail.batch(torch.randn(1, 3, 32, 32, device=DEV), metas, [torch.zeros(5, 5, dtype=torch.uint8)])
    with pytest.raises(RuntimeError):
        process_batch_reverse(torch.randn(1, 3, 32, 32), metas)          # CPU tensor: no CPU path
