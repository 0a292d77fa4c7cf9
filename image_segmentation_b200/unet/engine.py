"""How the U-Net (unet/unet.py:67-105) is laid out on the generic launch-plan engine (``image_segmentation_b200/engine.py``).

The reference runs ``unet.forward`` as ~90 ATen calls and lets autograd replay ~150 more.  Here the network is described
once per (batch, resolution, precision): five encoder levels whose second activation is written straight into the first
half of that level's concat buffer (and max-pooled in the same pass), four decoder levels whose ConvTranspose epilogue
writes the up-sampled map into the second half (``torch.cat([x1, up(x2)], 1)`` at unet/unet.py:63 never runs: skip FIRST,
up-sampled map SECOND), and the 1x1 head fused with the last BatchNorm + ReLU.
"""
from __future__ import annotations

import torch

from ..engine import Act, Engine, NetPlan

ENC_CH = (64, 128, 256, 512, 1024)


def double_conv(plan: NetPlan, name: str, block, src: Act, out=None, pool: bool = False, end_block: bool = False):
    """``DoubleConvReLU`` (unet/unet.py:13-25): (conv3x3 -> BatchNorm -> ReLU) x 2; ``block`` holds the parameters."""
    seq = block.doubleConvReLU
    l1 = plan.conv_bn_relu(name + ".c1", seq[0], seq[1], src, end_block=end_block)
    l2 = plan.conv_bn_relu(name + ".c2", seq[3], seq[4], l1.out, out=out, pool=pool)
    return l1, l2


def build_unet(plan: NetPlan, x: torch.Tensor):
    m = plan.model
    n, din, h, w = x.shape
    if h % 16 or w % 16:
        # the reference fails in torch.cat (unet/unet.py:63) for such sizes
        raise ValueError(f"U-Net input height/width must be multiples of 16, got {h}x{w}")
    widths = [m.scale * c for c in ENC_CH]
    src = plan.image_input(din, h, w)
    cats = [plan.cat(h >> l, w >> l, [widths[l], widths[l]], name=f"cat{l + 1}") for l in range(4)]
    enc_blocks = [m.down1] + [getattr(m, f"down{i}").maxpool_doubleConv[1] for i in range(2, 6)]
    prev = src
    for l in range(5):
        l1, l2 = double_conv(plan, f"down{l + 1}", enc_blocks[l], prev, out=cats[l].parts[0] if l < 4 else None,
                             pool=(l < 4), end_block=True)
        prev = l2.pooled if l < 4 else l2.out
    for i in range(4):
        up = getattr(m, f"up{i + 1}")
        l = 3 - i
        plan.conv_transpose(f"up{i + 1}.upsample", up.upsample, prev, out=cats[l].parts[1], end_block=True)
        l1, l2 = double_conv(plan, f"up{i + 1}", up.doubleConv, cats[l])
        prev = l2.out
    plan.head_1x1("output", m.output, prev)


def _check_input(model):
    def check(x):
        if x.dim() != 4 or x.shape[1] != model.din:
            raise RuntimeError(f"expected input [N,{model.din},H,W], got {tuple(x.shape)}")
    return check


def UNetEngine(model) -> Engine:
    """Plan cache + entry point used by ``unet.forward``."""
    return Engine(model, build_unet, _check_input(model))
