"""Drop-in for the reference's ``unet/unet.py``: same class names, constructor signatures, sub-module
names and ``state_dict`` keys (136 tensors), same default initialisation and RNG consumption order --
but ``unet.forward`` runs the B200-native engine (``engine.py``) instead of chaining ATen ops.

The nn.Conv2d / nn.BatchNorm2d / nn.ConvTranspose2d objects are kept purely as *parameter holders*
(so ``torch.manual_seed(s); unet(3, 3)`` is bit-identical to the reference, checkpoints load unchanged
and any optimizer works); their own ``forward`` methods are never called on the hot path.  The blocks
(``DoubleConvReLU``, ``Down``, ``Up``) stay callable on their own, as in the reference (unet/unet.py:24,44,62).

Reference: unet/unet.py:4-25 (DoubleConvReLU), :28-45 (Down), :47-64 (Up), :67-105 (unet).

Extra, non-reference attributes of ``unet``:
  ``precision``  "bf16" (default; tcgen05 kernels, fp32 accumulation) or "fp32" (CUDA-core parity tier)
  ``conv_algo``  "auto" | "simt" | "tc"  -- which contraction kernels run (debug / cross-check knob)
"""
from torch import nn

from ..engine import Engine, EngineModule, NhwcOutput
from .engine import UNetEngine


def _conv_bn_relu_pair(din, dout):
    layers = []
    for cin in (din, dout):
        layers += [nn.Conv2d(cin, dout, kernel_size=3, padding=1), nn.BatchNorm2d(dout), nn.ReLU()]
    return nn.Sequential(*layers)


class DoubleConvReLU(EngineModule):
    """(conv3x3 -> BatchNorm -> ReLU) x 2 (unet/unet.py:4-25).

    Inside ``unet`` the block is a parameter holder (the whole network runs as ONE fused pass); called on its own it runs
    as a small launch plan of its own: NCHW fp32 in -> NHWC -> the same fused kernels -> NCHW fp32 out, differentiable
    w.r.t. parameters and input.  Channel counts on the tensor-core tier: multiples of 64 (or an image-like input with
    at most 7 channels, which takes the first-layer im2col path and receives no input gradient)."""

    def __init__(self, din, dout):
        super().__init__()
        self.doubleConvReLU = _conv_bn_relu_pair(din, dout)

    def _build(self, plan, x):
        from .engine import double_conv
        n, c, h, w = x.shape
        src = plan.image_input(c, h, w) if 9 * c <= 64 else plan.nchw_input(0, x)
        _, l2 = double_conv(plan, "block", self, src)
        plan.add(NhwcOutput(plan, "block.out", l2.out))

    def forward(self, x):
        if self._engine is None:
            self._engine = Engine(self, self._build)
        return self._engine.run(x)


class Down(nn.Module):
    """MaxPool2d(2,2) then DoubleConvReLU (unet/unet.py:28-45).  Called on its own: torch's MaxPool2d, then the block's
    own launch plan (inside ``unet`` the pooling is fused into the producing BatchNorm-apply pass instead)."""

    def __init__(self, din, dout):
        super().__init__()
        self.maxpool_doubleConv = nn.Sequential(nn.MaxPool2d(kernel_size=2, stride=2), DoubleConvReLU(din, dout))

    def forward(self, x):
        return self.maxpool_doubleConv(x)


class Up(EngineModule):
    """ConvTranspose2d(k2,s2) on x2 + cat([x1, up(x2)], 1) + DoubleConvReLU (unet/unet.py:47-64).  Called on its own it is
    one launch plan: x1 lands in the first half of the concat buffer, the ConvTranspose epilogue writes the second."""

    def __init__(self, din, dout):
        super().__init__()
        self.upsample = nn.ConvTranspose2d(din, dout, kernel_size=2, stride=2)
        self.doubleConv = DoubleConvReLU(din, dout)

    def _build(self, plan, x1, x2):
        from .engine import double_conv
        n, c1, h, w = x1.shape
        if x2.shape[2] * 2 != h or x2.shape[3] * 2 != w:
            # the reference fails in torch.cat (unet/unet.py:63) for such sizes
            raise RuntimeError(f"Sizes of tensors must match: x1 {tuple(x1.shape)}, up(x2) from {tuple(x2.shape)}")
        cat = plan.cat(h, w, [c1, self.upsample.out_channels], name="cat")
        plan.nchw_input(0, x1, out=cat.parts[0])
        a2 = plan.nchw_input(1, x2)
        plan.conv_transpose("upsample", self.upsample, a2, out=cat.parts[1])
        _, l2 = double_conv(plan, "doubleConv", self.doubleConv, cat)
        plan.add(NhwcOutput(plan, "out", l2.out))

    def forward(self, x1, x2):
        if self._engine is None:
            self._engine = Engine(self, self._build)
        return self._engine.run(x1, x2)


class unet(nn.Module):
    """U-Net: 5 encoder levels (64..1024 channels), 4 decoder levels, 1x1 classifier head.

    ``forward(x)``: x is [N, din, H, W] fp32 on a CUDA device (H, W multiples of 16) -> logits
    [N, dout, H, W] fp32, differentiable w.r.t. every parameter.
    """

    def __init__(self, din, dout):
        super().__init__()
        self.scale = 1
        self.din, self.dout = din, dout
        widths = [self.scale * c for c in (64, 128, 256, 512, 1024)]
        self.down1 = DoubleConvReLU(din, widths[0])
        for i in range(1, 5):
            setattr(self, f"down{i + 1}", Down(widths[i - 1], widths[i]))
        for i in range(4):
            setattr(self, f"up{i + 1}", Up(widths[4 - i], widths[3 - i]))
        self.output = nn.Conv2d(widths[0], dout, kernel_size=1)
        self.precision = "bf16"
        self.conv_algo = "auto"
        self._engine = None
        self._train_graph = None

    def forward(self, x):
        if self._engine is None:
            self._engine = UNetEngine(self)
        return self._engine.run(x)

    def _apply(self, fn, *args, **kwargs):
        # .to()/.cuda()/.double() move or retype parameters: cached device buffers are then stale
        self._engine = None
        self._train_graph = None      # a captured step replays into the dropped engine's buffers (utils/training.py)
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        # device-side caches and process-group hooks never travel with a pickled / deep-copied model
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_train_graph"] = None
        for name in ("_bucket_hook", "_backward_done_hook", "_grad_scale", "_bucket_per_segment"):
            state.pop(name, None)
        return state
