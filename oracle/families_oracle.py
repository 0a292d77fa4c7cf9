"""CPU oracle for the other model families of the reference (TEST INFRASTRUCTURE ONLY).

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
may import it; nothing under ``image_segmentation_b200/`` does.  It restates, in functional form over a flat
``state_dict`` (so that it runs on any device / dtype the tensors live in), what the reference computes in

* ``autoencoder/autoencoder.py``   EncoderBlock :15-33, DecoderBlockWithSkips :69-93 (cat([up, skip]) :91),
                                    DecoderBlockNoSkips :128-147, ReconstructionAutoencoder :182-203 (Conv3x3 + Sigmoid
                                    :188-191), SegmentationAutoencoder :283-306
* ``clip/clipunet.py``             DecoderBlock :80-105 (ConvTranspose | 1x1 skip conv + bilinear resize, cat([x, skip])
                                    :102), UNetDecoder :121-146, output layer :183-188; tokens -> maps :46-63
* ``prompt_based/prompt.py``       probability composition :36-56
* ``utils/weighted_loss.py``       WeightedMemoryEfficientDiceLossPrompt :170-273, WeightedDiceNLLLoss :276-343

The arithmetic lives in PyTorch (unpinned third-party dependency of the reference; pin = this image's torch 2.11.0), so
the restatement uses ``torch.nn.functional`` primitives plus closed forms.  Parity pin: ``tests/golden/
make_golden_families.py`` runs the UNMODIFIED reference and stores its outputs (``autoencoder.npz``, ``clip.npz``,
``prompt.npz``); ``tests/test_oracle_families.py`` checks this file against them.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _bn_relu(sd, prefix, z, training, stats_out=None):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        mean = z.mean(dim=(0, 2, 3))
        var = z.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            stats_out[prefix] = (mean.detach(), z.var(dim=(0, 2, 3), unbiased=True).detach())
    else:
        mean, var = sd[prefix + ".running_mean"].to(z.dtype), sd[prefix + ".running_var"].to(z.dtype)
    xhat = (z - mean[None, :, None, None]) * torch.rsqrt(var + BN_EPS)[None, :, None, None]
    return torch.relu(xhat * w[None, :, None, None] + b[None, :, None, None])


def _conv_bn_relu(sd, conv, bn, x, training, stats_out=None):
    z = F.conv2d(x, sd[conv + ".weight"], sd.get(conv + ".bias"), padding=1)
    return _bn_relu(sd, bn, z, training, stats_out)


def _encoder(sd, pre, x, training, stats_out=None):
    """Encoder.forward (autoencoder.py:50-54): returns (bottleneck, skip3, skip2, skip1)."""
    skips = []
    for i in (1, 2, 3):
        p = f"{pre}encoderPart{i}"
        x = _conv_bn_relu(sd, p + ".conv1", p + ".bn1", x, training, stats_out)
        x = _conv_bn_relu(sd, p + ".conv2", p + ".bn2", x, training, stats_out)
        skips.append(x)
        x = F.max_pool2d(x, 2, 2)
    return x, skips[2], skips[1], skips[0]


def _convs(sd, p, x, training, stats_out=None):
    x = _conv_bn_relu(sd, p + ".0", p + ".1", x, training, stats_out)
    return _conv_bn_relu(sd, p + ".3", p + ".4", x, training, stats_out)


def reconstruction_forward(sd: Dict[str, torch.Tensor], x, training=True, stats_out=None):
    """ReconstructionAutoencoder.forward (autoencoder.py:193-203)."""
    h, _, _, _ = _encoder(sd, "encoder.", x, training, stats_out)
    for i in (1, 2, 3):
        p = f"decoder.decoderBlock{i}"
        h = F.conv_transpose2d(h, sd[p + ".up.weight"], sd[p + ".up.bias"], stride=2)
        h = _convs(sd, p + ".convs", h, training, stats_out)
    return torch.sigmoid(F.conv2d(h, sd["decoderOut.0.weight"], sd["decoderOut.0.bias"], padding=1))


def segmentation_ae_forward(sd: Dict[str, torch.Tensor], x, training=True, stats_out=None):
    """SegmentationAutoencoder.forward (autoencoder.py:296-306); cat order [up, skip] (:91)."""
    h, s3, s2, s1 = _encoder(sd, "encoder.encoder.", x, training, stats_out)
    for i, skip in ((1, s3), (2, s2), (3, s1)):
        p = f"decoder.decoderBlock{i}"
        up = F.conv_transpose2d(h, sd[p + ".up.weight"], sd[p + ".up.bias"], stride=2)
        h = _convs(sd, p + ".convs", torch.cat([up, skip], dim=1), training, stats_out)
    return F.conv2d(h, sd["finalConv.weight"], sd["finalConv.bias"])


def tokens_to_map(t: torch.Tensor, grid: int):
    """clip/clipunet.py:46-50: drop CLS, [N, grid*grid, C] -> [N, C, grid, grid]."""
    n, _, c = t.shape
    return t[:, 1:, :].reshape(n, grid, grid, c).permute(0, 3, 1, 2).contiguous()


def clip_decoder_forward(sd: Dict[str, torch.Tensor], tokens: List[torch.Tensor], grid: int, training=True, stats_out=None):
    """UNetDecoder.forward + output layer (clip/clipunet.py:139-146,183-188).  ``tokens`` = [last_hidden_state] +
    [hidden_states[i] for i in skip_indices]; ``sd`` holds the ``decoder.*`` and ``output_layer.*`` tensors."""
    x = tokens_to_map(tokens[0], grid)
    skips = [tokens_to_map(t, grid) for t in tokens[1:]]
    x = F.conv2d(x, sd["decoder.init_conv.weight"], sd["decoder.init_conv.bias"])
    nblocks = len({k.split(".")[2] for k in sd if k.startswith("decoder.decoder_blocks.")})
    for bi, skip in zip(range(nblocks), reversed(skips)):
        p = f"decoder.decoder_blocks.{bi}"
        x = F.conv_transpose2d(x, sd[p + ".upsample.weight"], sd[p + ".upsample.bias"], stride=2)
        skip = F.conv2d(skip, sd[p + ".skip_conv.weight"], sd[p + ".skip_conv.bias"])
        if skip.shape[2:] != x.shape[2:]:
            skip = F.interpolate(skip, size=x.shape[2:], mode="bilinear", align_corners=False)
        x = _convs(sd, p + ".conv_block", torch.cat([x, skip], dim=1), training, stats_out)
    return F.conv2d(x, sd["output_layer.weight"], sd["output_layer.bias"])


def prompt_compose(clip_logit: torch.Tensor, mask_logit: torch.Tensor):
    """prompt_based/prompt.py:36-56 in closed form: [1 - m, m p0 + m p3, m p1, m p2]."""
    p = torch.softmax(clip_logit, dim=1)
    m = torch.sigmoid(mask_logit)
    sel = m * p
    return torch.cat([1.0 - m, sel[:, 0:1] + sel[:, 3:4], sel[:, 1:2], sel[:, 2:3]], dim=1)


def dice_nll_loss(probs: torch.Tensor, target: torch.Tensor, dice_weight=1.0, nll_weight=1.0, ignore_index: Optional[int] = None,
                  class_weights: Optional[torch.Tensor] = None, smooth_dice=1e-5, nll_eps: Optional[float] = 1e-9):
    """WeightedDiceNLLLoss(apply_softmax=False, nll_nonlin=log(x + eps)) (utils/weighted_loss.py:276-343) in closed form.
    ``nll_eps=None`` = nll_nonlin None (NLLLoss on the raw input).  Differentiable w.r.t. ``probs``."""
    n, c, h, w = probs.shape
    oh = F.one_hot(target.long(), c).permute(0, 3, 1, 2).to(probs.dtype)
    inter = (probs * oh).sum(dim=(0, 2, 3))
    den = probs.sum(dim=(0, 2, 3)) + oh.sum(dim=(0, 2, 3))
    dc = (2.0 * inter + smooth_dice) / torch.clip(den + smooth_dice, 1e-8)
    valid = torch.ones(c, dtype=torch.bool)
    if ignore_index is not None and 0 <= ignore_index < c:
        valid[ignore_index] = False
    if class_weights is not None:
        cw = class_weights.to(probs.dtype)
        dice = (dc[valid] * cw[valid]).sum() / cw[valid].sum().clamp(min=1e-8)
    else:
        dice = dc[valid].mean()
    z = torch.log(probs + nll_eps) if nll_eps is not None else probs
    picked = z.gather(1, target.long().unsqueeze(1)).squeeze(1)
    wy = (class_weights.to(probs.dtype)[target.long()] if class_weights is not None else torch.ones_like(picked))
    if ignore_index is not None:
        wy = wy * (target != ignore_index).to(probs.dtype)
    nll = -(wy * picked).sum() / wy.sum()
    return dice_weight * (-dice) + nll_weight * nll
