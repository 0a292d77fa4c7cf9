"""`start()` (utils/training.py:453-617) end to end on the GPU against a golden run of the reference's own start():
two epochs of train_loop + eval_loop on tiny deterministic loaders, checkpoint files and keys, resume."""
import contextlib
import importlib.util
import io
import json
import os
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from image_segmentation_b200.unet.unet import unet  # noqa: E402
from image_segmentation_b200.utils.MetricsHistory import MetricsHistory  # noqa: E402
from image_segmentation_b200.utils.training import start  # noqa: E402
from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss  # noqa: E402

DEV = "cuda"
CLASS_W4 = [0.2046795970925636, 1.0271954434416883, 1.2293222812780409, 1.5388026781877073]
HERE = os.path.dirname(os.path.abspath(__file__))


def _fixture():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.start_fixture()


def _numbers(text, label):
    return [float(v) for v in re.findall(re.escape(label) + r"\s*:?\s*(-?\d+\.\d+)", text)]


def test_start_two_epochs_matches_reference_and_resumes(golden, tmp_path):
    g = golden["start"]
    train, val = _fixture()
    torch.manual_seed(0)
    m = unet(3, 4)
    m.precision = "fp32"
    opt = torch.optim.AdamW(m.parameters(), weight_decay=0.01)
    loss_fn = WeightedDiceCELoss(smooth_dice=1, class_weights=torch.tensor(CLASS_W4), ignore_index=3)
    agg = MetricsHistory(4, 3)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        start(str(tmp_path), "ck.pt", m, opt, train, val, 1, torch.device(DEV), loss_fn, loss_fn, 32, None, agg, True, True,
              4, 3, 2)
    ours, ref = buf.getvalue(), str(g["stdout"])
    np.testing.assert_allclose(_numbers(ours, "Training Avg loss (per effective batch)"),
                               _numbers(ref, "Training Avg loss (per effective batch)"), atol=2e-3)
    np.testing.assert_allclose(_numbers(ours, "Average Loss (Original Size)"), _numbers(ref, "Average Loss (Original Size)"),
                               atol=3e-3)
    # metrics depend on per-pixel argmax decisions of a barely trained model: a handful of near-ties may flip
    np.testing.assert_allclose(agg.get_mean_iou_history(), g["miou_history"], atol=5e-3)
    np.testing.assert_allclose(agg.get_mean_dice_history(), g["dice_history"], atol=5e-3)
    assert ours.count("Saving model...") == ref.count("Saving model...")
    files = sorted(os.listdir(tmp_path)) + sorted("metrics/" + f for f in os.listdir(tmp_path / "metrics"))
    assert files == json.loads(str(g["files"]))
    ck = torch.load(tmp_path / "ck.pt", weights_only=True)
    assert sorted(ck.keys()) == json.loads(str(g["ck_keys"])) and ck["epoch"] == int(g["ck_epoch"])
    np.testing.assert_allclose([ck["best_dev_dice"], ck["best_dev_miou"], ck["best_dev_loss"]], g["ck_best"], atol=5e-3)
    assert list(ck["model_state_dict"].keys()) == list(m.state_dict().keys())
    # resume: a fresh model + optimizer pick the checkpoint up and run epoch 3 only
    torch.manual_seed(1)
    m2 = unet(3, 4)
    m2.precision = "fp32"
    opt2 = torch.optim.AdamW(m2.parameters(), weight_decay=0.01)
    buf2 = io.StringIO()
    with contextlib.redirect_stdout(buf2):
        start(str(tmp_path), "ck.pt", m2, opt2, train, val, 1, torch.device(DEV), loss_fn, loss_fn, 32, None, MetricsHistory(4, 3),
              True, True, 4, 3, 3)
    out2 = buf2.getvalue()
    assert "Resuming training from epoch 3" in out2 and "Epoch 3" in out2 and "Epoch 2\n" not in out2
    assert opt2.state_dict()["state"][0]["step"] > 6          # optimizer state restored (6 steps) and advanced
