"""CPU-only checks of the drop-in boundary: module tree / state_dict / init stream, the C ABI library
(loads, exports every symbol include/unetk.h declares), and loud failure without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from image_segmentation_b200 import _lib as L
from image_segmentation_b200.unet.unet import unet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _digest(t):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()] + t[:4].tolist() + t[-4:].tolist()
                    if t.numel() >= 4 else [t.sum().item()] + t.tolist())


@pytest.mark.parametrize("din,dout", [(3, 3), (3, 4), (4, 1)])
def test_state_dict_and_init_match_reference(golden, din, dout):
    g = golden["init"]
    tag = f"{din}{dout}"
    torch.manual_seed(0)
    m = unet(din, dout)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g[f"keys_{tag}"])                       # 136 tensors, same names/order
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g[f"shapes_{tag}"])
    dig = np.stack([np.resize(_digest(v), 11) for v in sd.values()])
    np.testing.assert_array_equal(dig, g[f"digest_{tag}"])                 # same RNG stream -> bit-identical init
    assert len(list(m.parameters())) == 82
    # checkpoints round-trip through the reference's format (utils/training.py:506-509)
    m2 = unet(din, dout)
    m2.load_state_dict(sd)
    assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), sd.values()))


def test_header_symbols_are_exported():
    header = open(os.path.join(ROOT, "include", "unetk.h")).read()
    declared = set(re.findall(r"\b(unetk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    if not os.path.isfile(L.LIB_PATH):
        from image_segmentation_b200 import _build
        _build.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.unetk_version() >= 100


def test_struct_sizes_match_header():
    # POD structs are mirrored by hand in _lib.py: guard against drift with the sizes nvcc/gcc would produce
    assert ctypes.sizeof(L.Tensor) == 32
    assert ctypes.sizeof(L.ConvArgs) == 32 + 8 + 32 + 8 + 24 + 32 + 5 * 8
    assert ctypes.sizeof(L.WgradArgs) == 32 + 32 + 8 + 8 + 8 + 8
    assert ctypes.sizeof(L.BnBwdArgs) == 3 * 32 + 5 * 8 + 32 + 24
    assert ctypes.sizeof(L.BnFinalizeArgs) == 8 * 2 + 8 + 4 + 4 + 6 * 8 + 8 + 4 * 8
    assert ctypes.sizeof(L.DiceCeArgs) == 8 * 2 + 16 + 8 + 8 + 8 + 16 + 8 * 6 + 8
    assert ctypes.sizeof(L.HeadBnBwdArgs) == 32 + 8 + 8 + 8 + 4 * 8 + 8 + 32 + 4 * 8
    assert ctypes.sizeof(L.EvalImage) == 32
    assert ctypes.sizeof(L.EvalArgs) == 136      # static_assert'ed on the C side (csrc/eval.cu)
    # ... and against what the compiled library itself reports
    lib = L.lib()
    lib.unetk_struct_size.restype = ctypes.c_int32
    for which, cls in enumerate((L.Tensor, L.ConvArgs, L.WgradArgs, L.BnFinalizeArgs, L.BnBwdArgs, L.WJob, L.DiceCeArgs,
                                 L.HeadBnBwdArgs, L.EvalImage, L.EvalArgs)):
        assert lib.unetk_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.unetk_struct_size(99) == -1
    assert L.query_workspace(L.WS_DICE_ACCUM, 3) == 11 * 8 and L.query_workspace(L.WS_WGRAD, 1, 64, 128) == 64 * 9 * 128 * 4
    assert L.query_workspace(L.WS_POOL_IDX, 2, 8, 8, 64) == 2 * 4 * 4 * 8 * 2


def test_no_cpu_fallback():
    from image_segmentation_b200.utils.MetricsHistory import MetricsHistory
    from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss
    with pytest.raises(RuntimeError, match="CUDA"):
        unet(3, 3)(torch.zeros(1, 3, 16, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        WeightedDiceCELoss()(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        MetricsHistory(3).accumulate(torch.zeros(3, 4, 4), torch.zeros(4, 4, dtype=torch.long))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "image_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)


def test_batch_helpers_identity_at_training_resolution():
    from image_segmentation_b200.utils.utils import process_batch_forward, process_batch_reverse
    x = torch.rand(3, 3, 32, 32)
    out, metas = process_batch_forward(x, target_size=32)
    assert out is x and len(metas) == 3
    small = [torch.rand(3, 20, 30), torch.rand(3, 32, 16)]
    out, metas = process_batch_forward(small, target_size=32)
    assert out.shape == (2, 3, 32, 32)
    assert [m["original_size"] for m in metas] == [(20, 30), (32, 16)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        process_batch_reverse(out, metas)            # the reverse transform is a CUDA kernel, there is no CPU path


def _run_bench(*args, timeout=600):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(root, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=root)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): one JSON line with the contract's keys,
    timed on host cores, never touching the product."""
    import json
    r = _run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "unet256_train_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1 and d["value"] > 0
    assert "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_product_arm_fails_loudly_without_a_gpu():
    r = _run_bench("--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "needs a CUDA device" in r.stderr and "{\"metric\"" not in r.stdout
