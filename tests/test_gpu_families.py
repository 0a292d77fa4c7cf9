"""GPU parity of the other model families against the golden vectors of the UNMODIFIED
reference (tests/golden/autoencoder.npz, clip.npz, prompt.npz) and the CPU oracle (oracle/families_oracle.py).

Tolerances as for the U-Net (north_star): fp32 tier rel 1e-4 on activations / loss, gradients next to the reference's own
fp32-vs-fp64 deviation; bf16 tier rel-L2 5e-2 end to end (per-kernel 2e-2 is held in test_gpu_ops.py /
test_gpu_layers_b64.py); integer / closed-form kernels to float rounding.
"""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from image_segmentation_b200 import _lib as L  # noqa: E402
from image_segmentation_b200.utils.synthetic import make_batch  # noqa: E402
from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss, WeightedDiceNLLLoss  # noqa: E402
from oracle import families_oracle as FO  # noqa: E402

DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))
CLASS_W4 = [0.2046795970925636, 1.0271954434416883, 1.2293222812780409, 1.5388026781877073]


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden_families", os.path.join(HERE, "golden", "make_golden_families.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()


def rel_max(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def check_grads(model, g, tag_f32, tag_f64, factor=3.0, floor=2e-4, skip=()):
    """Every parameter gradient norm within `factor` x the reference's own fp32-vs-fp64 deviation (+ floor)."""
    names = list(g[f"grad_names_{tag_f64}"])
    n64, n32 = g[f"grad_norms_{tag_f64}"], g[f"grad_norms_{tag_f32}"]
    gd = dict(model.named_parameters())
    for k, a64, a32 in zip(names, n64, n32):
        p = gd[k]
        if np.isnan(a64):
            assert p.grad is None, f"{k}: frozen in the reference but has a gradient here"
            continue
        if k in skip or (".conv" in k and k.endswith(".bias")):
            continue
        got = p.grad.double().norm().item()
        ref_noise = abs(a32 - a64)
        assert abs(got - a64) <= factor * ref_noise + floor * max(a64, 1e-12) + 1e-9, (k, got, a64, a32)


def check_full_grads(model, g, tag, tol):
    gd = dict(model.named_parameters())
    mg = _mg()
    for key in g.files:
        if key.startswith(f"grad_{tag}:"):
            k = key.split(":", 1)[1]
            assert rel_max(gd[k].grad, g[key]) < tol, (k, rel_max(gd[k].grad, g[key]))
        elif key.startswith(f"gradsample_{tag}:"):
            k = key.split(":", 1)[1]
            got = gd[k].grad.flatten()[::mg.sample_stride(gd[k].numel())]
            assert rel_max(got, g[key]) < tol, (k, rel_max(got, g[key]))


# =====================================================================================================================
# N2 autoencoder family (autoencoder/autoencoder.py)
# =====================================================================================================================
def _recon(precision):
    from image_segmentation_b200.autoencoder.autoencoder import ReconstructionAutoencoder
    torch.manual_seed(0)
    m = ReconstructionAutoencoder(3)
    m.precision = precision
    return m.to(DEV).train()


def test_reconstruction_autoencoder_fp32_tier_matches_reference_golden(golden):
    g = golden["autoencoder"]
    x, _ = make_batch(2, 32, 32, 3, 4, seed=77, labels="learnable")
    m = _recon("fp32")
    out = m(x.to(DEV))
    loss = torch.nn.MSELoss()(out, x.to(DEV))
    loss.backward()
    assert rel_max(out, g["recon_out_f64"]) < 1e-4
    assert abs(loss.item() - float(g["recon_loss_f64"])) < 1e-4 * float(g["recon_loss_f64"])
    check_grads(m, g, "recon_f32", "recon_f64")
    check_full_grads(m, g, "recon_f64", 5e-2)      # element-wise: the reference's own fp32 run deviates by ~1e-2 (ReLU / pool flips)
    sd = m.state_dict()
    for key in g.files:
        if key.startswith("recon_buf_f64:"):
            np.testing.assert_allclose(sd[key.split(":", 1)[1]].cpu().numpy(), g[key], rtol=1e-4, atol=1e-6)
    m.eval()
    with torch.no_grad():
        assert rel_max(m(x.to(DEV)), g["recon_eval_f64"]) < 1e-4


def test_reconstruction_autoencoder_bf16_tier(golden):
    g = golden["autoencoder"]
    x, _ = make_batch(2, 32, 32, 3, 4, seed=77, labels="learnable")
    m = _recon("bf16")
    out = m(x.to(DEV))
    loss = torch.nn.MSELoss()(out, x.to(DEV))
    loss.backward()
    assert rel_l2(out, g["recon_out_f64"]) < 5e-2
    assert abs(loss.item() - float(g["recon_loss_f64"])) < 2e-2 * float(g["recon_loss_f64"])
    for p in m.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()
    # the same step on the CUDA-core fp32 tier: the two tiers must agree on the reconstruction-layer gradients
    m2 = _recon("fp32")
    loss2 = torch.nn.MSELoss()(m2(x.to(DEV)), x.to(DEV))
    loss2.backward()
    assert rel_l2(m.decoderOut[0].bias.grad, m2.decoderOut[0].bias.grad) < 0.1
    assert rel_l2(m.decoderOut[0].weight.grad, m2.decoderOut[0].weight.grad) < 0.15


@pytest.mark.parametrize("frozen", [True, False], ids=["frozen", "trainable"])
def test_segmentation_autoencoder_matches_reference_golden(golden, frozen, tmp_path, capsys):
    from image_segmentation_b200.autoencoder.autoencoder import ReconstructionAutoencoder, SegmentationAutoencoder
    g = golden["autoencoder"]
    x, y = make_batch(2, 32, 32, 3, 4, seed=77, labels="learnable")
    # the reference checkpoint = the reconstruction model after one fp64 training-mode step (no optimizer step): same
    # weights as the seeded init, running statistics moved once -- reproduced here with the fp32 tier
    rec = _recon("fp32")
    torch.nn.MSELoss()(rec(x.to(DEV)), x.to(DEV)).backward()
    ck = tmp_path / "recon.pt"
    torch.save({"model_state_dict": rec.state_dict()}, ck)
    which = "frozen" if frozen else "train"
    for precision, tol in (("fp32", 1e-4), ("bf16", 5e-2)):
        torch.manual_seed(1)
        s = SegmentationAutoencoder(3, 64, 4, pretrained_encoder_path=str(ck), freeze_encoder=frozen)
        s.precision = precision
        s = s.to(DEV).train()
        logits = s(x.to(DEV))
        loss = WeightedDiceCELoss(smooth_dice=1, class_weights=torch.tensor(CLASS_W4))(logits, y.squeeze(1).to(DEV))
        loss.backward()
        ref = g[f"seg_{which}_f64_logits"]
        err = rel_max(logits, ref) if precision == "fp32" else rel_l2(logits, ref)
        assert err < tol, (precision, err)
        assert abs(loss.item() - float(g[f"seg_{which}_f64_loss"])) < max(tol, 1e-4) * 2
        if precision == "fp32":
            check_grads(s, g, f"seg_{which}_f32", f"seg_{which}_f64")
            check_full_grads(s, g, f"seg_{which}_f64", 5e-2)
            np.testing.assert_allclose(s.state_dict()["encoder.encoder.encoderPart1.bn1.running_mean"].cpu().numpy(),
                                       g[f"seg_{which}_f64_buf:encoder.encoder.encoderPart1.bn1.running_mean"], rtol=1e-4, atol=1e-6)
        else:
            for k, p in s.named_parameters():
                assert (p.grad is None) == (frozen and k.startswith("encoder.")), k
    capsys.readouterr()


def test_train_reconstruction_loop_matches_reference_curve(golden):
    from image_segmentation_b200.utils.training import evalReconstruction, trainReconstruction
    g = golden["autoencoder"]
    m = _recon("fp32")
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    batches = [make_batch(2, 32, 32, 3, 4, seed=500 + i) for i in range(3)]
    curve = [trainReconstruction(batches, m, torch.nn.MSELoss(), opt, 1) for _ in range(4)]
    np.testing.assert_allclose(curve, g["recon_curve"], rtol=0, atol=2e-3)
    # evaluation at original (ragged) sizes runs and returns finite numbers
    val = [([torch.rand(3, 40, 28), torch.rand(3, 32, 32)], None), ([torch.rand(3, 25, 50)], None)]
    a, b = evalReconstruction(val, m, torch.nn.MSELoss(), 32)
    assert np.isfinite(a) and np.isfinite(b)


def test_autoencoder_at_training_resolution_bf16_vs_oracle():
    """256x256 (the size the reference trains at), batch 2, bf16 tensor-core tier against the fp32 oracle on the GPU."""
    from image_segmentation_b200.autoencoder.autoencoder import SegmentationAutoencoder
    torch.manual_seed(3)
    s = SegmentationAutoencoder(3, 64, 4, freeze_encoder=False).to(DEV).train()
    x, y = make_batch(2, 256, 256, 3, 4, seed=5)
    logits = s(x.to(DEV))
    sd = {k: v.detach() for k, v in s.state_dict().items()}
    # the oracle sees the state BEFORE this forward's running-stat update only through train-mode statistics: unaffected
    ref = FO.segmentation_ae_forward(sd, x.to(DEV), training=True)
    assert rel_l2(logits, ref) < 5e-2
    agree = (logits.argmax(1) == ref.argmax(1)).float().mean().item()
    assert agree > 0.97, agree


# =====================================================================================================================
# N3 CLIP decoder (clip/clipunet.py:68-188)
# =====================================================================================================================
def _tiny_clip(precision):
    from image_segmentation_b200.clip.clipunet import ClipUNet
    mg = _mg()
    torch.manual_seed(0)
    m = ClipUNet(num_classes=4, decoder_channels=mg.TINY_DECODER, clip_vit=mg.tiny_vit())
    m.precision = precision
    # (the tiny fixture has 32-channel layers (64 // 2); the tensor-core kernels need multiples of 64 -- every layer of the
    # real decoder [1024,512,256,128,64] is -- so in its bf16 run UNETK_ALGO_AUTO sends those layers to the CUDA-core
    # kernels on bf16 tensors; the tcgen05 path at the real geometry: test_clip_decoder_real_geometry_bf16_vs_oracle)
    return m.to(DEV).train()


def test_clip_decoder_fp32_tier_matches_reference_golden(golden):
    g = golden["clip"]
    m = _tiny_clip("fp32")
    toks = [torch.from_numpy(g[f"tokens_{i}"]).to(DEV) for i in range(5)]
    m.encoder.tokens = lambda x: toks                 # isolate the decoder: feed the reference's own tokens
    x = torch.from_numpy(g["x"]).to(DEV)
    y = torch.from_numpy(g["y"]).to(DEV)
    logits = m(x)
    loss = WeightedDiceCELoss(smooth_dice=1, class_weights=torch.tensor(CLASS_W4))(logits, y)
    loss.backward()
    assert rel_max(logits, g["logits_f32"]) < 1e-4
    assert abs(loss.item() - float(g["loss_f32"])) < 1e-4
    # the golden fp64 run used fp64 tokens; gradients are compared with the fp32 reference run on identical tokens
    names = list(g["grad_names_f32"])
    gd = dict(m.named_parameters())
    for k, ref_norm in zip(names, g["grad_norms_f32"]):
        if np.isnan(ref_norm):
            assert gd[k].grad is None, k
        else:
            got = gd[k].grad.double().norm().item()
            assert abs(got - ref_norm) <= 2e-2 * ref_norm + 1e-7, (k, got, ref_norm)
    check_full_grads(m, g, "f32", 5e-2)


def test_clip_unet_end_to_end_with_its_own_vit(golden):
    """Whole ClipUNet (torch ViT + engine decoder) against the golden logits; train and eval mode; bf16 tier."""
    g = golden["clip"]
    x = torch.from_numpy(g["x"]).to(DEV)
    # (the ViT is torch's: its GPU kernels -- fused attention, different reduction orders -- differ from the CPU run that
    # produced the golden tokens by ~1e-4..1e-3 at the logits; the decoder alone is held to 1e-4 in the test above)
    m = _tiny_clip("fp32")
    assert rel_max(m(x), g["logits_f32"]) < 3e-3
    m.eval()
    with torch.no_grad():
        assert rel_max(m(x), g["logits_eval_f32"]) < 3e-3
    mb = _tiny_clip("bf16")
    assert rel_l2(mb(x), g["logits_f64"]) < 5e-2


def test_clip_decoder_real_geometry_bf16_vs_oracle():
    """ViT-B/16 geometry (14x14x768 tokens -> 224x224 logits; 28..224 maps are not multiples of the 16x8 pixel tile)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel

    from image_segmentation_b200.clip.clipunet import ClipUNet
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(patch_size=16, num_hidden_layers=1, hidden_size=768, num_attention_heads=12, intermediate_size=64)
    m = ClipUNet(clip_vit=CLIPVisionModel(cfg)).to(DEV).train()          # the ViT is a stand-in: tokens are injected below
    g = torch.Generator().manual_seed(1)
    toks = [torch.randn(2, 197, 768, generator=g).to(DEV) for _ in range(5)]
    m.encoder.tokens = lambda x: toks
    y = torch.randint(0, 4, (2, 224, 224), generator=g).to(DEV)
    logits = m(torch.zeros(2, 3, 224, 224, device=DEV))
    loss = WeightedDiceCELoss(smooth_dice=1)(logits, y)
    loss.backward()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items() if not k.startswith("encoder.")}
    ref = FO.clip_decoder_forward(sd, toks, grid=14, training=True)
    assert rel_l2(logits, ref) < 5e-2
    from oracle import loss_oracle
    lo = loss_oracle.dice_ce_loss(ref, y, smooth_dice=1.0, dtype=torch.float32)
    lo.backward()
    gd = dict(m.named_parameters())
    for k in ("decoder.init_conv.weight", "decoder.init_conv.bias", "decoder.decoder_blocks.0.skip_conv.weight",
              "decoder.decoder_blocks.3.skip_conv.bias", "decoder.decoder_blocks.2.upsample.weight", "output_layer.weight"):
        e = rel_l2(gd[k].grad, sd[k].grad)
        assert e < 0.2, (k, e)                                           # bf16 end-to-end gradients (the reference under autocast: 0.36)


def test_bilinear_up_kernels_against_torch():
    import torch.nn.functional as F
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2)):
        for (ih, iw, oh, ow) in ((14, 14, 28, 28), (14, 14, 224, 224), (4, 4, 8, 8), (5, 7, 13, 9)):
            src = torch.randn(2, ih, iw, 64, device=DEV).to(dt)
            cat = torch.zeros(2, oh, ow, 128, device=DEV, dtype=dt)
            dst = cat[..., 64:]
            L.bilinear_up(src, dst)
            ref = F.interpolate(src.float().permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
            assert rel_max(dst, ref) < tol
            assert float(cat[..., :64].abs().sum()) == 0.0
            # backward = transpose of the forward: <up(s), d> == <s, up^T(d)>
            d = torch.randn(2, oh, ow, 64, device=DEV).to(dt)
            sref = src.float().permute(0, 3, 1, 2).requires_grad_(True)
            F.interpolate(sref, size=(oh, ow), mode="bilinear", align_corners=False).backward(d.float().permute(0, 3, 1, 2))
            ds = torch.empty_like(src)
            L.bilinear_up(ds, d, backward=True)
            assert rel_max(ds, sref.grad.permute(0, 2, 3, 1)) < max(tol, 1e-5)


# =====================================================================================================================
# N4 prompt model (prompt_based/prompt.py, utils/weighted_loss.py:170-343)
# =====================================================================================================================
def test_prompt_compose_kernels_match_reference_golden(golden):
    from image_segmentation_b200.prompt_based.prompt import _PromptCompose
    g = golden["prompt"]
    clip = torch.from_numpy(g["compose_clip"]).to(DEV)
    mask = torch.from_numpy(g["compose_mask"]).to(DEV).requires_grad_(True)
    final = _PromptCompose.apply(clip, mask)
    np.testing.assert_allclose(final.detach().cpu().numpy(), g["compose_final"], rtol=0, atol=1e-6)
    final.backward(torch.from_numpy(g["compose_up"]).to(DEV))
    np.testing.assert_allclose(mask.grad.cpu().numpy(), g["compose_dmask"], rtol=0, atol=2e-6)


def test_dice_nll_loss_matches_reference_golden(golden):
    g = golden["prompt"]
    stable_log = lambda t: torch.log(t + 1e-9)  # noqa: E731
    for ci, kw in enumerate(json.loads(str(g["nll_cases"]))):
        kw = dict(kw)
        if "class_weights" in kw:
            kw["class_weights"] = torch.tensor(kw["class_weights"])
        fn = WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=stable_log, **kw)
        probs = torch.from_numpy(g[f"nll_probs_{ci}"]).to(DEV).requires_grad_(True)
        target = torch.from_numpy(g[f"nll_target_{ci}"]).to(DEV)
        loss = fn(probs, target)
        loss.backward()
        assert abs(loss.item() - float(g[f"nll_loss_{ci}"])) < 2e-6, ci
        assert rel_max(probs.grad, g[f"nll_grad_{ci}"]) < 1e-5, ci
        loss4 = fn(probs.detach(), target.unsqueeze(1))                     # [N,1,H,W] targets
        assert abs(loss4.item() - loss.item()) < 1e-7
    with pytest.raises(NotImplementedError):
        WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=torch.sqrt)
    with pytest.raises(ValueError):
        WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=stable_log)(probs.detach(), target[0])


def test_dice_nll_loss_large_random_vs_oracle():
    g = torch.Generator().manual_seed(3)
    probs = torch.softmax(torch.randn(8, 4, 224, 224, generator=g) * 2, dim=1)
    target = torch.randint(0, 4, (8, 224, 224), generator=g)
    w = torch.tensor(CLASS_W4)
    fn = WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=lambda t: torch.log(t + 1e-9), smooth_dice=1, class_weights=w, ignore_index=3)
    pd = probs.to(DEV).requires_grad_(True)
    loss = fn(pd, target.to(DEV))
    loss.backward()
    pc = probs.double().requires_grad_(True)
    ref = FO.dice_nll_loss(pc, target, smooth_dice=1.0, class_weights=w.double(), ignore_index=3)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 2e-6
    assert rel_max(pd.grad, pc.grad) < 1e-5


def test_prompt_model_matches_reference_golden(golden):
    from image_segmentation_b200.clip.clipunet import ClipUNet
    from image_segmentation_b200.prompt_based.prompt import PromptModel
    mg = _mg()
    g = golden["prompt"]
    x, heat, y = (torch.from_numpy(g[k]).to(DEV) for k in ("pm_x", "pm_heat", "pm_y"))
    # fp32 tier: 2e-3 instead of 1e-4 because the frozen ViT runs on torch's GPU kernels here and on the CPU in the golden run
    for precision, tol in (("fp32", 2e-3), ("bf16", 5e-2)):
        torch.manual_seed(0)
        pm = PromptModel(clip=ClipUNet(num_classes=4, decoder_channels=mg.TINY_DECODER, clip_vit=mg.tiny_vit()))
        pm.clip.precision = pm.mask.precision = precision
        pm = pm.to(DEV).train()
        probs = pm(x, heat)
        fn = WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=lambda t: torch.log(t + 1e-9), smooth_dice=1,
                                 class_weights=torch.tensor(CLASS_W4))
        loss = fn(probs, y)
        loss.backward()
        err = rel_max(probs, g["pm_probs"]) if precision == "fp32" else rel_l2(probs, g["pm_probs"])
        assert err < tol, (precision, err)
        assert abs(loss.item() - float(g["pm_loss"])) < (5e-4 if precision == "fp32" else 2e-2)
        assert torch.allclose(probs.sum(1), torch.ones_like(probs[:, 0]), atol=1e-5)
        assert all(p.grad is None for p in pm.clip.parameters())
        if precision == "fp32":
            names = list(g["grad_names_pm_mask"])
            gd = dict(pm.mask.named_parameters())
            for k, ref_norm in zip(names, g["grad_norms_pm_mask"]):
                if k.endswith(".bias") and "doubleConvReLU" in k and k.split(".")[-2] in ("0", "3"):
                    continue                                   # conv bias in front of BatchNorm: exact zero here, noise in the reference
                got = gd[k].grad.double().norm().item()
                assert abs(got - ref_norm) <= 3e-2 * ref_norm + 1e-7, (k, got, ref_norm)
            for key in ("output.weight", "output.bias"):
                assert rel_max(gd[key].grad, g[f"grad_pm_mask:{key}"]) < 2e-3, key


def test_start_prompt_trains_evaluates_and_checkpoints(tmp_path, capsys):
    """start_prompt / train_loop_prompt / eval_loop_prompt (utils/training.py:153-199,242-450) on the tiny fixture: the mask
    network trains (loss falls), the frozen CLIP branch does not move, the checkpoint carries the history object and there
    is no MO_ weights copy; a second call resumes from the checkpoint."""
    from image_segmentation_b200.clip.clipunet import ClipUNet
    from image_segmentation_b200.prompt_based.prompt import PromptModel
    from image_segmentation_b200.utils.MetricsHistory import MetricsHistory
    from image_segmentation_b200.utils.training import start_prompt
    mg = _mg()
    torch.manual_seed(0)
    pm = PromptModel(clip=ClipUNet(num_classes=4, decoder_channels=mg.TINY_DECODER, clip_vit=mg.tiny_vit()))
    pm.clip.precision = pm.mask.precision = "fp32"
    pm = pm.to(DEV)
    clip_before = {k: v.detach().clone() for k, v in pm.clip.named_parameters()}
    g = torch.Generator().manual_seed(4)
    train = []
    for _ in range(3):
        x = torch.rand(2, 3, 64, 64, generator=g)
        heat = torch.rand(2, 1, 64, 64, generator=g)
        y = (x.mean(1, keepdim=True) * 3.999).floor().to(torch.uint8)
        train.append((x, heat, y))
    val = [([torch.rand(3, 50, 70, generator=g)], [torch.rand(1, 50, 70, generator=g)], [torch.randint(0, 4, (50, 70), generator=g)])]
    opt = torch.optim.AdamW([p for p in pm.parameters() if p.requires_grad], lr=1e-3)
    fn = WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=lambda t: torch.log(t + 1e-9), smooth_dice=1,
                             class_weights=torch.tensor(CLASS_W4))
    start_prompt(str(tmp_path), "pm.pt", pm, opt, train, val, 1, torch.device(DEV), fn, fn, 64, None, MetricsHistory(4, 3), True, True,
                 4, 3, 2)
    out = capsys.readouterr().out
    losses = [float(v) for v in __import__("re").findall(r"Training Avg loss \(per effective batch\):\s*(-?\d+\.\d+)", out)]
    assert len(losses) == 2 and losses[1] < losses[0]
    for k, v in pm.clip.named_parameters():
        assert torch.equal(v, clip_before[k]), k
    files = sorted(os.listdir(tmp_path))
    assert "pm.pt" in files and not any(f.startswith("MO_") for f in files)
    ck = torch.load(tmp_path / "pm.pt", weights_only=False)
    assert "history" in ck and ck["epoch"] in (1, 2)
    start_prompt(str(tmp_path), "pm.pt", pm, opt, train, val, 1, torch.device(DEV), fn, fn, 64, None, None, True, True, 4, 3, 3)
    out2 = capsys.readouterr().out
    assert f"Resuming training from epoch {ck['epoch'] + 1}" in out2 and " -> Metrics History loaded." in out2
