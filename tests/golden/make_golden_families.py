"""Golden vectors for the other model families from the UNMODIFIED reference:
``python tests/golden/make_golden_families.py`` (authoring container only; /root/reference is imported through
``oracle/ref_shim.py``, nothing in it is edited).

  autoencoder.npz  ReconstructionAutoencoder / SegmentationAutoencoder (autoencoder/autoencoder.py): init digests, one
                   fwd+bwd step in fp32 and fp64 (outputs, loss, gradient norms, a few full gradients, running statistics,
                   eval-mode output), frozen and trainable encoder, trainReconstruction loss curve (utils/training.py:123-151)
  clip.npz         ClipUNet decoder (clip/clipunet.py:68-188) on the tokens of a tiny random-init CLIP ViT: tokens in,
                   logits / loss / gradients out
  prompt.npz       PromptModel composition (prompt_based/prompt.py:33-56) and WeightedDiceNLLLoss
                   (utils/weighted_loss.py:276-343) forward / backward
"""
from __future__ import annotations

import contextlib
import io
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
CLASS_W4 = [0.2046795970925636, 1.0271954434416883, 1.2293222812780409, 1.5388026781877073]
TINY_VIT = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=10, num_attention_heads=2, image_size=64,
                patch_size=16)
TINY_DECODER = [256, 128, 64, 64, 64]


def tensor_digest(t: torch.Tensor):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()] + t[:4].tolist()
                    + t[-4:].tolist() if t.numel() >= 4 else [t.sum().item()] + t.tolist())


def init_digest(m, out, tag):
    sd = m.state_dict()
    out[f"keys_{tag}"] = np.array(list(sd.keys()))
    out[f"shapes_{tag}"] = np.array([str(tuple(v.shape)) for v in sd.values()])
    out[f"digest_{tag}"] = np.stack([np.resize(tensor_digest(v), 11) for v in sd.values()])


def sample_stride(numel: int) -> int:
    return max(1, -(-numel // 4096)) | 1          # odd stride: walks through every position of the inner dimensions


def grads_of(m, out, tag, full=()):
    names, norms = [], []
    gd = dict(m.named_parameters())
    for k, p in gd.items():
        names.append(k)
        norms.append(float("nan") if p.grad is None else p.grad.double().norm().item())
    out[f"grad_names_{tag}"] = np.array(names)
    out[f"grad_norms_{tag}"] = np.array(norms)
    for k in full:
        g = gd[k].grad.detach()
        if g.numel() <= 8192:
            out[f"grad_{tag}:{k}"] = g.numpy()
        else:
            # large tensors: every stride-th element (<= 4096 values) keeps the fixture small and still pins the layout
            out[f"gradsample_{tag}:{k}"] = g.flatten()[::sample_stride(g.numel())].numpy()


def gen_autoencoder(ref):
    ae = ref.autoencoder_mod
    out = {}
    x, y = make_batch(2, 32, 32, 3, 4, seed=77, labels="learnable")
    for dt, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
        # ---- reconstruction pretraining model ----
        torch.manual_seed(0)
        m = ae.ReconstructionAutoencoder(3).to(dt).train()
        if dn == "f32":
            init_digest(m, out, "recon")
        rec = m(x.to(dt))
        loss = torch.nn.MSELoss()(rec, x.to(dt))
        loss.backward()
        out[f"recon_out_{dn}"] = rec.detach().numpy()
        out[f"recon_loss_{dn}"] = np.array(loss.item())
        grads_of(m, out, f"recon_{dn}", ("encoder.encoderPart1.conv1.weight", "encoder.encoderPart3.bn2.weight",
                                         "decoder.decoderBlock1.up.bias", "decoder.decoderBlock3.convs.3.weight",
                                         "decoderOut.0.weight", "decoderOut.0.bias"))
        sd = m.state_dict()
        for k in ("encoder.encoderPart1.bn1.running_mean", "encoder.encoderPart1.bn1.running_var",
                  "decoder.decoderBlock2.convs.4.running_var"):
            out[f"recon_buf_{dn}:{k}"] = sd[k].numpy()
        m.eval()
        with torch.no_grad():
            out[f"recon_eval_{dn}"] = m(x.to(dt)).numpy()
        ck = os.path.join(ROOT, "build", "golden_recon.pt")
        os.makedirs(os.path.dirname(ck), exist_ok=True)
        torch.save({"model_state_dict": m.state_dict()}, ck)
        # ---- segmentation model re-using the encoder (frozen / trainable) ----
        for frozen in (True, False):
            tag = f"seg_{'frozen' if frozen else 'train'}_{dn}"
            torch.manual_seed(1)
            with contextlib.redirect_stdout(io.StringIO()) as buf:
                s = ae.SegmentationAutoencoder(3, 64, 4, pretrained_encoder_path=ck, freeze_encoder=frozen).to(dt).train()
            if dn == "f32":
                out[f"seg_stdout_{'frozen' if frozen else 'train'}"] = np.array(buf.getvalue())
                if frozen:
                    init_digest(s, out, "seg")
            w = torch.tensor(CLASS_W4, dtype=dt)
            loss_fn = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w)
            logits = s(x.to(dt))
            loss = loss_fn(logits, y.squeeze(1))
            loss.backward()
            out[f"{tag}_logits"] = logits.detach().numpy()
            out[f"{tag}_loss"] = np.array(loss.item())
            full = ["decoder.decoderBlock1.convs.0.weight", "decoder.decoderBlock3.up.weight", "finalConv.weight", "finalConv.bias"]
            if not frozen:
                full.append("encoder.encoder.encoderPart2.conv2.weight")
            grads_of(s, out, tag, full)
            out[f"{tag}_buf:encoder.encoder.encoderPart1.bn1.running_mean"] = s.state_dict()["encoder.encoder.encoderPart1.bn1.running_mean"].numpy()
    # ---- trainReconstruction through the reference's own loop (module-global device = cpu here) ----
    tm = ref.training_mod
    torch.manual_seed(0)
    m = ae.ReconstructionAutoencoder(3)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    batches = [make_batch(2, 32, 32, 3, 4, seed=500 + i) for i in range(3)]
    curve = []
    for _ in range(4):
        with contextlib.redirect_stdout(io.StringIO()):
            curve.append(tm.trainReconstruction(batches, m, torch.nn.MSELoss(), opt, 1))
    out["recon_curve"] = np.array(curve)
    out["cfg"] = np.array(json.dumps(dict(n=2, hw=32, seed=77, class_weights=CLASS_W4)))
    np.savez_compressed(os.path.join(OUT, "autoencoder.npz"), **out)


def tiny_vit(seed=3):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        return CLIPVisionModel(CLIPVisionConfig(**TINY_VIT)).eval()
    finally:
        torch.random.set_rng_state(state)


@contextlib.contextmanager
def patched_from_pretrained(vit):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    a, b = CLIPVisionConfig.from_pretrained, CLIPVisionModel.from_pretrained
    CLIPVisionConfig.from_pretrained = staticmethod(lambda *x, **k: vit.config)
    CLIPVisionModel.from_pretrained = staticmethod(lambda *x, **k: vit)
    try:
        yield
    finally:
        CLIPVisionConfig.from_pretrained, CLIPVisionModel.from_pretrained = a, b


def gen_clip(ref):
    cm = ref.clip_mod
    out = {}
    vit = tiny_vit()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 64, 64, generator=g)
    y = torch.randint(0, 4, (2, 64, 64), generator=g)
    for dt, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
        torch.manual_seed(0)
        with patched_from_pretrained(vit.to(dt)):
            m = cm.ClipUNet(num_classes=4, decoder_channels=TINY_DECODER).to(dt).train()
        if dn == "f32":
            dec_sd = {k: v for k, v in m.state_dict().items() if not k.startswith("encoder.")}
            out["keys"] = np.array(list(dec_sd.keys()))
            out["digest"] = np.stack([np.resize(tensor_digest(v), 11) for v in dec_sd.values()])
            with torch.no_grad():
                o = m.encoder.clip_vit(pixel_values=x, output_hidden_states=True)
            toks = [o.last_hidden_state] + [o.hidden_states[i] for i in m.encoder.skip_indices]
            for i, t in enumerate(toks):
                out[f"tokens_{i}"] = t.numpy()
        w = torch.tensor(CLASS_W4, dtype=dt)
        logits = m(x.to(dt))
        loss = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w)(logits, y)
        loss.backward()
        out[f"logits_{dn}"] = logits.detach().numpy()
        out[f"loss_{dn}"] = np.array(loss.item())
        grads_of(m, out, dn, ("decoder.init_conv.weight", "decoder.init_conv.bias", "decoder.decoder_blocks.0.skip_conv.weight",
                              "decoder.decoder_blocks.3.skip_conv.bias", "decoder.decoder_blocks.1.upsample.weight",
                              "decoder.decoder_blocks.2.conv_block.0.weight", "output_layer.weight"))
        m.eval()
        with torch.no_grad():
            out[f"logits_eval_{dn}"] = m(x.to(dt)).numpy()
    out["x"] = x.numpy()
    out["y"] = y.numpy()
    out["cfg"] = np.array(json.dumps(dict(vit=TINY_VIT, decoder_channels=TINY_DECODER, vit_seed=3, class_weights=CLASS_W4)))
    np.savez_compressed(os.path.join(OUT, "clip.npz"), **out)


def gen_prompt(ref):
    out = {}
    g = torch.Generator().manual_seed(9)
    # ---- composition (prompt_based/prompt.py:36-56), isolated from the networks ----
    clip_logit = torch.randn(2, 4, 12, 10, generator=g) * 2
    mask_logit = (torch.randn(2, 1, 12, 10, generator=g) * 2).requires_grad_(True)
    clip_prob = torch.softmax(clip_logit, dim=1)
    mask_prob = torch.sigmoid(mask_logit)
    final = torch.empty_like(clip_prob)
    sel = mask_prob * clip_prob
    final[:, 1:4] = sel[:, 0:3]
    final[:, 0:1] = 1.0 - mask_prob
    final[:, 1:2] += sel[:, 3:4]
    up = torch.randn(2, 4, 12, 10, generator=g)
    final.backward(up)
    out["compose_clip"] = clip_logit.numpy()
    out["compose_mask"] = mask_logit.detach().numpy()
    out["compose_final"] = final.detach().numpy()
    out["compose_up"] = up.numpy()
    out["compose_dmask"] = mask_logit.grad.numpy()
    # ---- WeightedDiceNLLLoss on probabilities (prompt_based/prompt.ipynb:68-70) ----
    stable_log = lambda t: torch.log(t + 1e-9)  # noqa: E731
    cases = []
    for ci, kw in enumerate([dict(smooth_dice=1.0), dict(smooth_dice=1.0, class_weights=CLASS_W4),
                             dict(class_weights=CLASS_W4, ignore_index=3), dict(dice_weight=0.6, nll_weight=1.7, ignore_index=0)]):
        probs = torch.softmax(torch.randn(3, 4, 9, 11, generator=g) * 2, dim=1).requires_grad_(True)
        target = torch.randint(0, 4, (3, 9, 11), generator=g)
        kw_t = dict(kw)
        if "class_weights" in kw_t:
            kw_t["class_weights"] = torch.tensor(kw_t["class_weights"])
        fn = ref.loss_mod.WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=stable_log, **kw_t)
        loss = fn(probs, target)
        loss.backward()
        out[f"nll_probs_{ci}"] = probs.detach().numpy()
        out[f"nll_target_{ci}"] = target.numpy()
        out[f"nll_loss_{ci}"] = np.array(loss.item())
        out[f"nll_grad_{ci}"] = probs.grad.numpy()
        cases.append(kw)
    out["nll_cases"] = np.array(json.dumps(cases))
    # ---- whole PromptModel on the tiny ViT: probabilities + gradient norms of the mask network ----
    pm = ref.prompt_mod
    vit = tiny_vit()
    torch.manual_seed(0)
    with patched_from_pretrained(vit):
        orig = pm.ClipUNet
        pm.ClipUNet = lambda: orig(num_classes=4, decoder_channels=TINY_DECODER)     # PromptModel() builds ClipUNet() itself
        try:
            model = pm.PromptModel().train()
        finally:
            pm.ClipUNet = orig
    x = torch.rand(2, 3, 64, 64, generator=g)
    heat = torch.rand(2, 1, 64, 64, generator=g)
    y = torch.randint(0, 4, (2, 64, 64), generator=g)
    fn = ref.loss_mod.WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=stable_log, smooth_dice=1, class_weights=torch.tensor(CLASS_W4))
    probs = model(x, heat)
    loss = fn(probs, y)
    loss.backward()
    out["pm_x"], out["pm_heat"], out["pm_y"] = x.numpy(), heat.numpy(), y.numpy()
    out["pm_probs"] = probs.detach().numpy()
    out["pm_loss"] = np.array(loss.item())
    grads_of(model.mask, out, "pm_mask", ("output.weight", "output.bias", "down1.doubleConvReLU.0.weight"))
    out["pm_trainable"] = np.array(sorted(k for k, p in model.named_parameters() if p.requires_grad))
    np.savez_compressed(os.path.join(OUT, "prompt.npz"), **out)


def gen_unet_bf16_noise(ref):
    """The reference's OWN deviation under bf16 autocast (CPU) from its fp64 run, per parameter gradient and for the
    logits, on the inputs of unet_step.npz (2x3x32x32, seed 1234) and at 2x3x256x256 (seed 7): the yardstick for the
    end-to-end bf16 tolerances of tests/test_gpu_unet.py."""
    out = {}
    for tag, hw, seed in (("32", 32, 1234), ("256", 256, 7)):
        x, y = make_batch(2, hw, hw, 3, 3, seed=seed)
        res = {}
        for mode in ("f64", "bf16"):
            torch.manual_seed(0)
            m = ref.unet(3, 3).train()
            w = torch.tensor(CLASS_W4[:3])
            loss_fn = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w)
            if mode == "f64":
                m = m.double()
                logits = m(x.double())
                loss = ref.WeightedDiceCELoss(smooth_dice=1, class_weights=w.double())(logits, y.squeeze(1))
            else:
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    logits = m(x)
                loss = loss_fn(logits.float(), y.squeeze(1))
            loss.backward()
            res[mode] = (logits.detach().double(), {k: p.grad.detach().double() for k, p in m.named_parameters()}, loss.item())
        l64, g64, loss64 = res["f64"]
        l16, g16, loss16 = res["bf16"]
        out[f"names_{tag}"] = np.array(list(g64.keys()))
        out[f"grad_rel_l2_{tag}"] = np.array([((g16[k] - g64[k]).norm() / g64[k].norm().clamp(min=1e-30)).item() for k in g64])
        out[f"logits_rel_l2_{tag}"] = np.array(((l16 - l64).norm() / l64.norm()).item())
        out[f"loss_{tag}"] = np.array([loss64, loss16])
    np.savez_compressed(os.path.join(OUT, "unet_bf16_noise.npz"), **out)


def main():
    ref = ref_shim.load()
    torch.set_num_threads(8)
    gen_unet_bf16_noise(ref)
    gen_autoencoder(ref)
    gen_clip(ref)
    gen_prompt(ref)
    print("family golden vectors written to", OUT)


if __name__ == "__main__":
    main()
