"""CPU oracle for the U-Net forward/backward of the reference (TEST INFRASTRUCTURE ONLY).

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product path (``image_segmentation_b200``) never imports anything under ``oracle/``.

It restates, in functional form over a flat ``state_dict``, what the reference computes in
``unet/unet.py`` (all citations are into ``/root/reference``):

* ``DoubleConvReLU``  unet/unet.py:13-25   conv3x3(p=1,bias) -> BN(train) -> ReLU, twice
* ``Down``            unet/unet.py:37-45   MaxPool2d(2,2) then DoubleConvReLU
* ``Up``              unet/unet.py:56-64   ConvTranspose2d(k2,s2) on x2; cat([x1, up(x2)],1); DoubleConvReLU
* ``unet``            unet/unet.py:76-105  5 encoder levels 64..1024, 4 decoder levels, 1x1 head

The arithmetic of the reference lives in a third-party dependency that is not vendored and
not pinned by the reference (PyTorch); the pin used here is the container's torch 2.11.0.
The restatement therefore uses ``torch.nn.functional`` primitives on CPU tensors (fp32 = the
reference's own precision, fp64 = ground truth) plus closed-form formulas where they are
short (BatchNorm, ConvTranspose as an einsum).

Parity pin: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference modules from
``/root/reference`` in this container and stores their outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this oracle against those vectors (and, when
``/root/reference`` is present, against the live reference).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

ENC_CH = (64, 128, 256, 512, 1024)
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# parameter naming / initialisation (mirrors the module tree built at unet/unet.py:80-91)
# --------------------------------------------------------------------------------------
def double_conv_prefixes() -> List[Tuple[str, int, int]]:
    """(state_dict prefix, cin, cout) of the 9 DoubleConvReLU blocks, in construction order."""
    out = [("down1.doubleConvReLU", None, 64)]
    for i, (ci, co) in enumerate(zip(ENC_CH[:-1], ENC_CH[1:]), start=2):
        out.append((f"down{i}.maxpool_doubleConv.1.doubleConvReLU", ci, co))
    return out


def _uniform_like_torch_conv(shape, fan_in, gen=None):
    # nn.Conv2d.reset_parameters: kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    # same floating-point operation order as torch.nn.init.kaiming_uniform_ so bounds match bit-for-bit
    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
    std = gain / math.sqrt(fan_in)
    bound = math.sqrt(3.0) * std
    return torch.empty(shape).uniform_(-bound, bound, generator=gen)


def init_state_dict(din: int, dout: int, gen=None) -> "OrderedDict[str, torch.Tensor]":
    """Random-init state_dict consuming the RNG stream in the same order as ``unet(din, dout)``.

    unet/unet.py:80-91 builds down1..down5, up1..up4, output; each nn.Conv2d /
    nn.ConvTranspose2d draws weight then bias (torch.nn.modules.conv._ConvNd.reset_parameters);
    BatchNorm2d draws nothing.  ConvTranspose2d's fan_in is computed from weight.size(1)*k*k,
    i.e. ``Cout*4`` for the ``[Cin, Cout, 2, 2]`` layout.
    """
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def conv(prefix, cin, cout, k):
        fan_in = cin * k * k
        sd[prefix + ".weight"] = _uniform_like_torch_conv((cout, cin, k, k), fan_in, gen)
        b = 1.0 / math.sqrt(fan_in)
        sd[prefix + ".bias"] = torch.empty(cout).uniform_(-b, b, generator=gen)

    def bn(prefix, c):
        sd[prefix + ".weight"] = torch.ones(c)
        sd[prefix + ".bias"] = torch.zeros(c)
        sd[prefix + ".running_mean"] = torch.zeros(c)
        sd[prefix + ".running_var"] = torch.ones(c)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def double_conv(prefix, cin, cout):
        conv(prefix + ".0", cin, cout, 3)
        bn(prefix + ".1", cout)
        conv(prefix + ".3", cout, cout, 3)
        bn(prefix + ".4", cout)

    double_conv("down1.doubleConvReLU", din, 64)
    for i, (ci, co) in enumerate(zip(ENC_CH[:-1], ENC_CH[1:]), start=2):
        double_conv(f"down{i}.maxpool_doubleConv.1.doubleConvReLU", ci, co)
    for i, (ci, co) in enumerate(zip(ENC_CH[:0:-1], ENC_CH[-2::-1]), start=1):
        # nn.ConvTranspose2d(ci, co, 2, 2): weight [ci, co, 2, 2]; fan_in = co*4
        fan_in = co * 4
        sd[f"up{i}.upsample.weight"] = _uniform_like_torch_conv((ci, co, 2, 2), fan_in, gen)
        b = 1.0 / math.sqrt(fan_in)
        sd[f"up{i}.upsample.bias"] = torch.empty(co).uniform_(-b, b, generator=gen)
        double_conv(f"up{i}.doubleConv.doubleConvReLU", ci, co)
    conv("output", 64, dout, 1)
    return sd


def param_names(sd) -> List[str]:
    """Names of the trainable tensors (everything except BN buffers), in state_dict order."""
    return [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


# --------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------
def batchnorm_train(z, gamma, beta, eps=BN_EPS):
    """Train-mode BatchNorm2d in closed form: biased variance for normalisation.

    Returns (y, mean, biased_var).  Matches torch.nn.functional.batch_norm(training=True)
    (unet/unet.py:17,20).
    """
    mean = z.mean(dim=(0, 2, 3))
    var = z.var(dim=(0, 2, 3), unbiased=False)
    xhat = (z - mean[None, :, None, None]) * torch.rsqrt(var + eps)[None, :, None, None]
    return xhat * gamma[None, :, None, None] + beta[None, :, None, None], mean, var


def conv_transpose_2x2(x, w, b):
    """ConvTranspose2d(k=2, s=2) as the einsum it is (unet/unet.py:59).

    out[n, co, 2i+a, 2j+b] = sum_ci x[n, ci, i, j] * w[ci, co, a, b] + bias[co]
    """
    n, _, h, wd = x.shape
    co = w.shape[1]
    y = torch.einsum("ncij,cdab->ndiajb", x, w).reshape(n, co, 2 * h, 2 * wd)
    return y + b[None, :, None, None]


def _double_conv(x, sd, prefix, training, new_buffers, native_ops=False):
    for idx_conv, idx_bn in (("0", "1"), ("3", "4")):
        z = F.conv2d(x, sd[f"{prefix}.{idx_conv}.weight"], sd[f"{prefix}.{idx_conv}.bias"], padding=1)
        g, b = sd[f"{prefix}.{idx_bn}.weight"], sd[f"{prefix}.{idx_bn}.bias"]
        rm, rv = sd[f"{prefix}.{idx_bn}.running_mean"], sd[f"{prefix}.{idx_bn}.running_var"]
        if native_ops:
            # exactly the ATen calls nn.BatchNorm2d makes (used by the CPU timing legs of bench.py so that the baseline
            # runs the reference's own operators, not the closed-form restatement below)
            if training:
                nbt = sd[f"{prefix}.{idx_bn}.num_batches_tracked"]
                y = F.batch_norm(z, rm, rv, g, b, True, BN_MOMENTUM, BN_EPS)     # updates rm / rv in place
                with torch.no_grad():
                    nbt += 1
            else:
                y = F.batch_norm(z, rm, rv, g, b, False, BN_MOMENTUM, BN_EPS)
            x = torch.relu(y)
            continue
        if training:
            y, mean, var = batchnorm_train(z, g, b)
            if new_buffers is not None:
                cnt = z.numel() // z.shape[1]
                with torch.no_grad():
                    new_buffers[f"{prefix}.{idx_bn}.running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
                    new_buffers[f"{prefix}.{idx_bn}.running_var"] = (
                        (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * (cnt / max(cnt - 1, 1)))
                    new_buffers[f"{prefix}.{idx_bn}.num_batches_tracked"] = sd[f"{prefix}.{idx_bn}.num_batches_tracked"] + 1
        else:
            scale = g * torch.rsqrt(rv + BN_EPS)
            y = (z - rm[None, :, None, None]) * scale[None, :, None, None] + b[None, :, None, None]
        x = torch.relu(y)
    return x


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True,
            new_buffers: Dict[str, torch.Tensor] | None = None,
            taps: Dict[str, torch.Tensor] | None = None, native_ops: bool = False) -> torch.Tensor:
    """unet.forward (unet/unet.py:93-105).  ``taps`` (optional) receives intermediate activations."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    skips = []
    h = tap("x1", _double_conv(x, sd, "down1.doubleConvReLU", training, new_buffers, native_ops))
    skips.append(h)
    for i in range(2, 6):
        h = F.max_pool2d(h, kernel_size=2, stride=2)
        h = tap(f"x{i}", _double_conv(h, sd, f"down{i}.maxpool_doubleConv.1.doubleConvReLU", training, new_buffers, native_ops))
        if i < 5:
            skips.append(h)
    for i in range(1, 5):
        if native_ops:
            up = F.conv_transpose2d(h, sd[f"up{i}.upsample.weight"], sd[f"up{i}.upsample.bias"], stride=2)
        else:
            up = conv_transpose_2x2(h, sd[f"up{i}.upsample.weight"], sd[f"up{i}.upsample.bias"])
        h = torch.cat([skips[4 - i], up], dim=1)           # skip FIRST (unet/unet.py:63)
        h = tap(f"u{i}", _double_conv(h, sd, f"up{i}.doubleConv.doubleConvReLU", training, new_buffers, native_ops))
    return F.conv2d(h, sd["output.weight"], sd["output.bias"])


# --------------------------------------------------------------------------------------
# training step built on autograd over the functional forward (the reference relies on the same
# autograd engine: utils/training.py:46-50)
# --------------------------------------------------------------------------------------
def loss_and_grads(sd, x, y, loss_fn, dtype=torch.float32, training=True):
    """Returns (loss, logits, {param: grad}, new_buffers) for one micro-batch."""
    names = param_names(sd)
    work = {k: (v.detach().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    for k in names:
        work[k].requires_grad_(True)
    new_buffers: Dict[str, torch.Tensor] = {}
    logits = forward(work, x.to(dtype), training=training, new_buffers=new_buffers)
    loss = loss_fn(logits, y)
    grads = torch.autograd.grad(loss, [work[k] for k in names])
    return loss.detach(), logits.detach(), dict(zip(names, grads)), new_buffers


class OracleUNet(torch.nn.Module):
    """nn.Module face of the oracle so that it can be handed to an optimizer / train loop.

    Holds the flat state_dict as parameters/buffers under the reference's names (dots replaced
    internally); used by bench.py's CPU baseline leg and by the loss-curve parity tests.
    """

    def __init__(self, din: int, dout: int, gen=None, native_ops: bool = False):
        super().__init__()
        self.native_ops = native_ops
        sd = init_state_dict(din, dout, gen)
        self._names = list(sd.keys())
        for k, v in sd.items():
            key = k.replace(".", "__")
            if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                self.register_buffer(key, v)
            else:
                self.register_parameter(key, torch.nn.Parameter(v))

    def flat(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, k.replace(".", "__")) for k in self._names}

    def reference_state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self.flat().items())

    def load_reference_state_dict(self, sd):
        with torch.no_grad():
            for k, v in sd.items():
                getattr(self, k.replace(".", "__")).copy_(v)

    def forward(self, x):
        sd = self.flat()
        if self.native_ops:
            return forward(sd, x, training=self.training, native_ops=True)
        new_buffers: Dict[str, torch.Tensor] = {}
        out = forward(sd, x, training=self.training, new_buffers=new_buffers)
        with torch.no_grad():
            for k, v in new_buffers.items():
                getattr(self, k.replace(".", "__")).copy_(v)
        return out
