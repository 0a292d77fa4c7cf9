"""Composable launch plans over libunetk.so: the forward/backward of a whole network as one autograd node.

The reference builds its networks from a handful of blocks -- conv3x3 + BatchNorm + ReLU (unet/unet.py:13-25,
autoencoder/autoencoder.py:15-33,74-81, clip/clipunet.py:86-93), MaxPool2d(2,2) (unet/unet.py:40,
autoencoder/autoencoder.py:23), ConvTranspose2d(k2,s2) + channel concatenation (unet/unet.py:59-63,
autoencoder/autoencoder.py:72,91, clip/clipunet.py:83,102), 1x1 convolutions (unet/unet.py:91, clip/clipunet.py:84,125),
a 3x3 convolution + Sigmoid (autoencoder/autoencoder.py:188-191) and bilinear up-sampling (clip/clipunet.py:99-100) --
and lets autograd replay ~150 ATen calls per step.  Here a model family describes its network ONCE to a ``NetPlan``
(``plan.conv_bn_relu(...)``, ``plan.conv_transpose(...)``, ``plan.head(...)`` ... in forward order); the plan owns
the NHWC device buffers, prepares every C-ABI call once (``_lib.Call``: argument structs with stable device pointers)
and derives the backward pass itself, including the fusions:

  * BatchNorm statistics in the producing conv's epilogue, BatchNorm apply + ReLU (+ MaxPool + 2-bit arg-max) in one
    bandwidth pass that writes straight into a concat slice (``torch.cat`` never runs);
  * BatchNorm-backward reductions fused into the data-gradient launch that produces the activation gradient, whenever
    that gradient has a single source (otherwise the stand-alone reduction kernel merges skip + pool gradients);
  * head (1x1, <= 4 classes) fused with the last block's BatchNorm apply (forward) and BatchNorm backward;
  * frozen sub-networks (``requires_grad=False``, autoencoder/autoencoder.py:257-260): no weight gradients, and no data
    gradients below the first trainable layer (a conv over a concat of [trainable, frozen] inputs computes only the
    trainable channel range of its data gradient);
  * parameter gradients land in ONE flat fp32 buffer in backward completion order (data-parallel buckets, parallel.py).

Host code is plumbing: PyTorch allocates, this file sequences launches on the current stream.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import _lib as L


def dtype_of(precision: str):
    if precision == "bf16":
        return torch.bfloat16
    if precision == "fp32":
        return torch.float32
    raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")


class Act:
    """An NHWC activation: a [N,H,W,C] view (possibly a channel slice of a concat buffer) plus, once a backward pass
    has been planned, the view its gradient is written to."""

    def __init__(self, t: torch.Tensor, needs_grad: bool, producer=None, name: str = ""):
        self.t, self.needs_grad, self.producer, self.name = t, needs_grad, producer, name
        self.consumers: list = []
        self.parent: Optional["Cat"] = None     # concat buffer this view is a slice of
        self.offset = 0
        self.grad: Optional[torch.Tensor] = None

    @property
    def c(self):
        return self.t.shape[3]


class Cat(Act):
    """A concat buffer [N,H,W,sum(C_i)] whose parts are written in place by their producers (zero-copy torch.cat)."""

    def __init__(self, t, parts_c: Sequence[int], name=""):
        self.parts: List[Act] = []
        super().__init__(t, False, None, name)
        off = 0
        for c in parts_c:
            a = Act(t[..., off:off + c], False, None, f"{name}[{off}:{off + c}]")
            a.parent, a.offset = self, off
            self.parts.append(a)
            off += c

    @property
    def needs_grad(self):
        return any(p.needs_grad for p in self.parts)

    @needs_grad.setter
    def needs_grad(self, value):      # derived from the parts
        pass


# =====================================================================================================================
# nodes
# =====================================================================================================================
class Node:
    name = ""
    end_block = False       # a data-parallel bucket / batched weight-gradient unpack may close after this node's backward
    can_fuse_reduce = False  # this node's data gradient may also accumulate its producer's BatchNorm-backward sums
    fuse_reduce_of = None
    src: Optional[Act] = None

    def params(self) -> List[torch.nn.Parameter]:
        return []

    # extension nodes (everything except the three core node types) implement:
    def outputs(self) -> List[Act]:
        return []

    def runs_backward(self) -> bool:
        return False

    def ws_size(self) -> int:
        return 0

    def forward_calls(self, plan: "NetPlan", training: bool) -> list:
        return []

    def backward_calls(self, plan: "NetPlan", ops: list, jobs_pending: list) -> int:
        """Append the backward launches to `ops`; returns how many trainable parameters became final."""
        return 0


class ConvBNReLU(Node):
    """conv3x3(p=1) -> BatchNorm2d -> ReLU [-> MaxPool2d(2,2)] (unet/unet.py:16-18, autoencoder/autoencoder.py:17-23)."""

    def __init__(self, name, conv, bn, src: Act, first: bool):
        self.name, self.conv, self.bn, self.src, self.first = name, conv, bn, src, first
        self.cin, self.cout = conv.in_channels, conv.out_channels
        self.z = self.out = self.pooled = None
        self.pool_idx = None
        self.fused_head: Optional["Head"] = None
        self.wf = self.wd = self.ws = None
        self.dz = None
        self.reduced = False          # BatchNorm-backward sums are accumulated by the launch that produces d(out)
        self.fuse_reduce_of: Optional["ConvBNReLU"] = None   # this node's dgrad accumulates that layer's sums
        self.can_fuse_reduce = not first

    def params(self):
        ps = [self.conv.weight]
        if self.conv.bias is not None:
            ps.append(self.conv.bias)
        return ps + [self.bn.weight, self.bn.bias]


class ConvT(Node):
    """ConvTranspose2d(k=2, s=2) writing into (a slice of) a concat buffer (unet/unet.py:59,63)."""

    def __init__(self, name, mod, src: Act, out: Act):
        self.name, self.mod, self.src, self.out = name, mod, src, out
        self.cin, self.cout = mod.in_channels, mod.out_channels
        self.wf = self.wd = self.ws = None
        self.fuse_reduce_of: Optional[ConvBNReLU] = None
        self.can_fuse_reduce = True

    def params(self):
        return [self.mod.weight] + ([self.mod.bias] if self.mod.bias is not None else [])


class Head(Node):
    """1x1 classifier producing NCHW fp32 logits (unet/unet.py:91, autoencoder/autoencoder.py:295)."""

    def __init__(self, name, conv, src: Act):
        self.name, self.conv, self.src = name, conv, src
        self.dout = conv.out_channels
        self.fused = False
        self.g_in = None

    def params(self):
        return [self.conv.weight] + ([self.conv.bias] if self.conv.bias is not None else [])



def _flat_rows(t: torch.Tensor) -> torch.Tensor:
    """A contiguous [N,H,W,C] tensor seen as [1, M/pw, pw, C] (M = N*H*W rows of a plain GEMM; 1x1 convolutions have no
    spatial structure, and tiles of 128 consecutive rows keep the tensor-core tiles full whatever H and W are)."""
    n, h, w, c = t.shape
    if not t.is_contiguous():
        return t
    m = n * h * w
    pw = next(p for p in (16, 8, 4, 2, 1) if m % p == 0)
    return t.view(1, m // pw, pw, c)


class ConvSigmoidOut(Node):
    """nn.Conv2d(C, dout, 3, padding=1) + nn.Sigmoid producing the NCHW fp32 reconstruction
    (autoencoder/autoencoder.py:188-191).  The contraction runs on the tensor cores with Cout zero-padded to 64; bias and
    sigmoid are one bandwidth pass that reads only the first 8 channels."""
    can_fuse_reduce = True
    PAD = 64

    def __init__(self, plan: "NetPlan", name, conv, src: Act):
        if conv.kernel_size != (3, 3) or conv.padding != (1, 1) or conv.out_channels > 8:
            raise NotImplementedError("reconstruction output: conv3x3 padding 1 with at most 8 output channels")
        self.name, self.conv, self.src = name, conv, src
        self.cin, self.dout = conv.in_channels, conv.out_channels
        n, h, w, _ = src.t.shape
        dt, dev = plan.dt, plan.device
        self.z = plan.act(h, w, self.PAD)
        self.wf = torch.zeros((self.PAD, 9, self.cin), dtype=dt, device=dev)
        self.wd = torch.zeros((self.cin, 9, self.PAD), dtype=dt, device=dev)
        self.dz = None
        self.ws = None
        plan.output = torch.empty((n, self.dout, h, w), dtype=torch.float32, device=dev)

        def pack():
            wt = self.conv.weight.detach()
            self.wf[:self.dout].copy_(wt.permute(0, 2, 3, 1).reshape(self.dout, 9, self.cin))
            self.wd[:, :, :self.dout].copy_(wt.flip(2, 3).permute(1, 2, 3, 0).reshape(self.cin, 9, self.dout))
        plan._extra_pack.append(pack)

    def params(self):
        return [self.conv.weight] + ([self.conv.bias] if self.conv.bias is not None else [])

    def runs_backward(self):
        return any(p.requires_grad for p in self.params()) or self.src.needs_grad

    def ws_size(self):
        return self.PAD * 9 * self.cin if self.conv.weight.requires_grad else 0

    def forward_calls(self, plan, training):
        return [L.prep_conv(self.src.t, self.wf, self.z, L.MODE_3X3, algo=plan.algo, label=self.name,
                            algo_flops=2 * self.z.shape[0] * self.z.shape[1] * self.z.shape[2] * 9 * self.cin * self.dout),
                L.prep_bias_sigmoid_fwd(self.z, self.conv.bias, self.dout, plan.output, label=self.name)]

    def backward_calls(self, plan, ops, jobs_pending):
        if self.dz is None:
            self.dz = torch.zeros_like(self.z)       # channels >= 8 stay zero forever; the kernel rewrites channels 0..7
        bias = self.conv.bias
        call = L.prep_bias_sigmoid_bwd(plan._dlogits_slot, plan.output, self.dout, self.dz, None, label=self.name)
        if bias is not None and bias.requires_grad:
            ops.append(_PatchedArg(call, 4, plan._goff(bias)))
        else:
            ops.append(call)
        dg = plan._dgrad_call(self, L.MODE_3X3, self.dz, self.name)
        if dg is not None:
            ops.append(dg)
        w = self.conv.weight
        if w.requires_grad:
            flops = 2 * self.z.shape[0] * self.z.shape[1] * self.z.shape[2] * 9 * self.cin * self.dout
            ops.append(plan.wgrad_call(self.dz, self.src.t, self.ws, 1, algo_flops=flops, label=self.name))
            o0 = plan._off_of[id(w)]

            def unpack(o0=o0, w=w):
                g = plan.flat_grad[o0:o0 + w.numel()].view(w.shape)
                g.copy_(self.ws.view(self.PAD, 3, 3, self.cin)[:self.dout].permute(0, 3, 1, 2))
            ops.append(_TorchCall(unpack))
        return len([p for p in self.params() if p.requires_grad])


class Conv1x1(Node):
    """nn.Conv2d(Cin, Cout, kernel_size=1) with bias, no normalisation (clip/clipunet.py:84,125): a plain GEMM over all
    pixels with the bias in the epilogue."""

    def __init__(self, plan: "NetPlan", name, conv, src: Act, out: Optional[Act] = None):
        if conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.padding != (0, 0):
            raise NotImplementedError("Conv1x1 node: kernel_size 1, stride 1, no padding")
        self.name, self.conv, self.src = name, conv, src
        self.cin, self.cout = conv.in_channels, conv.out_channels
        n, h, w, _ = src.t.shape
        if out is None:
            out = Act(plan.act(h, w, self.cout), False, None, name + ".out")
        out.producer = self
        out.needs_grad = any(p.requires_grad for p in self.params()) or src.needs_grad
        self.out = out
        self.wf = torch.empty((self.cout, self.cin), dtype=plan.dt, device=plan.device)
        self.wd = torch.empty((self.cin, self.cout), dtype=plan.dt, device=plan.device)

        def pack():
            w2 = self.conv.weight.detach().view(self.cout, self.cin)
            self.wf.copy_(w2)
            self.wd.copy_(w2.t())
        plan._extra_pack.append(pack)

    def params(self):
        return [self.conv.weight] + ([self.conv.bias] if self.conv.bias is not None else [])

    def outputs(self):
        return [self.out]

    def runs_backward(self):
        return self.out.needs_grad

    def forward_calls(self, plan, training):
        return [L.prep_conv(_flat_rows(self.src.t), self.wf, _flat_rows(self.out.t), L.MODE_1X1, bias=self.conv.bias, algo=plan.algo,
                            label=self.name)]

    def backward_calls(self, plan, ops, jobs_pending):
        dy = self.out.grad
        if self.src.needs_grad:
            ops.append(L.prep_conv(_flat_rows(dy), self.wd, _flat_rows(self.src.grad), L.MODE_1X1, algo=plan.algo, label=self.name))
        w, b = self.conv.weight, self.conv.bias
        if w.requires_grad:
            # dw[cout][1][cin] is the parameter layout itself: accumulate straight into the flat gradient buffer
            call = plan.wgrad_call(_flat_rows(dy), _flat_rows(self.src.t), plan.ws, 0, label=self.name)
            call.patch_ptr(call.keep[0], "dw", plan._goff(w))
            ops.append(call)
        if b is not None and b.requires_grad:
            ops.append(_PatchedArg(plan.channel_sum_call(dy, self.name), 1, plan._goff(b)))
        return len([p for p in self.params() if p.requires_grad])


class BilinearUp(Node):
    """F.interpolate(x, size, mode='bilinear', align_corners=False) (clip/clipunet.py:99-100) into a concat slice."""

    def __init__(self, name, src: Act, out: Act):
        self.name, self.src, self.out = name, src, out
        out.producer = self
        out.needs_grad = src.needs_grad

    def outputs(self):
        return [self.out]

    def runs_backward(self):
        return self.out.needs_grad

    def forward_calls(self, plan, training):
        return [L.prep_bilinear_up(self.src.t, self.out.t, label=self.name)]

    def backward_calls(self, plan, ops, jobs_pending):
        ops.append(L.prep_bilinear_up(self.src.grad, self.out.grad, backward=True, label=self.name))
        return 0


class NhwcOutput(Node):
    """The plan's result is an NHWC activation returned as NCHW fp32 (block-level API: DoubleConvReLU / Down / Up called on
    their own, unet/unet.py:24,44,62)."""

    def __init__(self, plan: "NetPlan", name, src: Act):
        self.name, self.src = name, src
        n, h, w, c = src.t.shape
        plan.output = torch.empty((n, c, h, w), dtype=torch.float32, device=plan.device)

    def runs_backward(self):
        return self.src.needs_grad

    def forward_calls(self, plan, training):
        return [L.prep_nhwc_to_nchw(self.src.t, plan.output, label=self.name)]

    def backward_calls(self, plan, ops, jobs_pending):
        ops.append(L.prep_nchw_to_nhwc(plan._dlogits_slot, self.src.grad, label=self.name))
        return 0


class NetPlan:
    """Device buffers + prepared launch sequence for one (batch, resolution, precision) problem of one network."""

    def __init__(self, model, n: int, precision: str, device):
        self.model, self.n, self.precision, self.device = model, n, precision, device
        self.dt = dtype_of(precision)
        self.busy = False
        self.generation = 0
        self.algo = L.ALGO_AUTO
        self.nodes: List[Node] = []
        self.layers: List[ConvBNReLU] = []
        self.cats: List[Cat] = []
        self.input_call_builders: list = []      # callables (x tensors) -> None, run before the node list
        self.feature_inputs: List[Act] = []      # NCHW feature-map inputs that may need a gradient (block-level API)
        self.output: Optional[torch.Tensor] = None
        self.head: Optional[Head] = None
        self._fwd_calls: Dict[bool, list] = {}
        self._bwd = None
        self._pack_versions = None
        self._pack_jobs = None
        self._pack_ptrs = None
        self._extra_pack: list = []              # python callables for packs the batched kernel does not cover
        self.training_pass = False
        self.flat_grad = None
        self.pair_first = False
        self.kpad = 0
        self.xcol = None
        # model.deterministic = True: weight gradients use the ordered two-pass split reduction (bit-reproducible runs)
        self.deterministic = bool(getattr(model, "deterministic", False)) or os.environ.get("UNETK_DETERMINISTIC", "0") == "1"
        self._det_sites: list = []
        self._partial = None
        self.fuse_head_enabled = os.environ.get("UNETK_FUSE_HEAD", "1") == "1"
        self.fuse_reduce_enabled = os.environ.get("UNETK_FUSE_BN_REDUCE", "1") == "1"

    # ---- builder API (called by the model families in forward order) -------------------------------------------------
    def act(self, h, w, c, n=None):
        return torch.empty((self.n if n is None else n, h, w, c), dtype=self.dt, device=self.device)

    def image_input(self, din: int, h: int, w: int) -> Act:
        """NCHW fp32 network input feeding a first conv with tiny Cin: the input is written as the im2col operand of
        that conv (K = 9*din padded), which then runs as a 1x1 contraction (see _lib / layout.cu)."""
        self.din, self.h, self.w = din, h, w
        self.kpad = ((9 * din + 63) // 64) * 64
        # pixel-pair form (bf16 tier): K padded to 32 per pixel, two horizontally adjacent pixels make one 64-wide GEMM row
        # and the block-diagonal weight [[W 0] [0 W]] (N = 128) writes both pixels' 64 channels = the NHWC pair.
        self.pair_first = (self.dt == torch.bfloat16 and w % 2 == 0 and 9 * din <= 32
                           and os.environ.get("UNETK_FIRST_PAIR", "1") == "1")
        if self.pair_first:
            self.kpad = 32
        self.xcol = self.act(h, w, self.kpad)
        a = Act(self.xcol, False, None, "input.im2col")
        a.is_image = True

        def run_input(inputs, stream):
            # X.to(device) + the first conv's implicit im2col (utils/training.py:45, unet/unet.py:16)
            L.prep_im2col3x3_first(inputs[0], self.xcol)(stream)
        self.input_call_builders.append(run_input)
        return a

    def cat(self, h, w, parts_c: Sequence[int], name="cat") -> Cat:
        c = Cat(self.act(h, w, sum(parts_c)), parts_c, name)
        self.cats.append(c)
        return c

    def conv_bn_relu(self, name, conv, bn, src: Act, out: Optional[Act] = None, pool: bool = False, end_block=False):
        """Returns the activated output Act (and sets ``node.pooled`` when ``pool``)."""
        first = bool(getattr(src, "is_image", False))
        node = ConvBNReLU(name, conv, bn, src, first)
        node.end_block = end_block
        n, h, w, _ = src.t.shape
        c = node.cout
        if conv.kernel_size != (3, 3) or conv.padding != (1, 1) or conv.stride != (1, 1) or conv.groups != 1:
            raise NotImplementedError("only conv3x3 stride 1 padding 1 is on the accelerated path")
        if not first and conv.in_channels != src.c:
            raise ValueError(f"{name}: conv expects {conv.in_channels} input channels, source has {src.c}")
        trainable = any(p.requires_grad for p in node.params())
        node.z = self.act(h, w, c)
        if out is None:
            out = Act(self.act(h, w, c), False, None, name + ".a")
        if tuple(out.t.shape) != (n, h, w, c):
            raise ValueError(f"{name}: output view has shape {tuple(out.t.shape)}, expected {(n, h, w, c)}")
        out.producer, out.needs_grad = node, trainable or src.needs_grad
        node.out = out
        if pool:
            if h % 2 or w % 2:
                raise ValueError(f"{name}: MaxPool2d(2,2) needs even height/width, got {h}x{w}")
            node.pooled = Act(self.act(h // 2, w // 2, c), out.needs_grad, node, name + ".pooled")
            node.pool_idx = torch.empty((n, h // 2, w // 2, c // 8), dtype=torch.int16, device=self.device)
        src.consumers.append(node)
        self.nodes.append(node)
        self.layers.append(node)
        return node

    def conv_transpose(self, name, mod, src: Act, out: Optional[Act] = None, end_block=False) -> ConvT:
        if mod.kernel_size != (2, 2) or mod.stride != (2, 2) or mod.padding != (0, 0):
            raise NotImplementedError("only ConvTranspose2d(kernel_size=2, stride=2) is on the accelerated path")
        n, h, w, _ = src.t.shape
        if out is None:
            out = Act(self.act(2 * h, 2 * w, mod.out_channels), False, None, name + ".out")
        node = ConvT(name, mod, src, out)
        node.end_block = end_block
        out.producer = node
        out.needs_grad = any(p.requires_grad for p in node.params()) or src.needs_grad
        src.consumers.append(node)
        self.nodes.append(node)
        return node

    def head_1x1(self, name, conv, src: Act) -> Head:
        node = Head(name, conv, src)
        p = src.producer
        node.fused = (self.fuse_head_enabled and isinstance(p, ConvBNReLU) and node.dout <= 4 and src.c == 64
                      and not src.consumers and p.pooled is None and src.parent is None)
        if node.fused:
            p.fused_head = node
        src.consumers.append(node)
        n, h, w, _ = src.t.shape
        self.output = torch.empty((n, node.dout, h, w), dtype=torch.float32, device=self.device)
        self.head = node
        self.nodes.append(node)
        return node

    def add(self, node: Node) -> Node:
        """Register an extension node (ConvSigmoidOut, Conv1x1, BilinearUp, NhwcOutput) in forward order."""
        if node.src is not None:
            node.src.consumers.append(node)
        self.nodes.append(node)
        return node

    def nchw_input(self, index: int, x: torch.Tensor, out: Optional[Act] = None) -> Act:
        """Input number `index` of the plan is an NCHW fp32 feature map (block-level API): converted to NHWC on the way in,
        its gradient converted back on the way out."""
        n, c, h, w = x.shape
        if out is None:
            out = Act(self.act(h, w, c), False, None, f"input{index}")
        out.needs_grad = bool(x.requires_grad)
        out.input_index = index
        self.feature_inputs.append(out)

        def run_input(inputs, stream, out=out, index=index):
            L.prep_nchw_to_nhwc(inputs[index], out.t)(stream)
        self.input_call_builders.append(run_input)
        return out

    def token_input(self, index: int, x: torch.Tensor, grid: int) -> Act:
        """Input number `index` is a ViT hidden state [N, 1 + grid*grid, C] (clip/clipunet.py:46-63): the patch tokens
        (CLS dropped) ARE the NHWC feature map [N, grid, grid, C]; only the dtype changes."""
        n, t, c = x.shape
        if t != 1 + grid * grid:
            raise ValueError(f"expected {1 + grid * grid} tokens, got {t}")
        if x.requires_grad:
            raise NotImplementedError("gradients into the ViT encoder (freeze_encoder=False) are not on the accelerated path")
        out = Act(self.act(grid, grid, c), False, None, f"tokens{index}")

        def run_input(inputs, stream, out=out, index=index):
            out.t.view(n, grid * grid, c).copy_(inputs[index][:, 1:, :])
        self.input_call_builders.append(run_input)
        return out

    # ---- finishing the plan -----------------------------------------------------------------------------------------
    def finish(self):
        """Per-channel scratch and operand packs, once every node is known."""
        dev, dt = self.device, self.dt
        tot_c = sum(l.cout for l in self.layers)
        # sum, sumsq, (+ 2 x 128 for the pixel-pair first layer, whose statistics arrive as two halves)
        self.acc64 = torch.zeros(2 * tot_c + 256, dtype=torch.float64, device=dev)
        self.first_stats = self.acc64[2 * tot_c:]
        self.vec32 = torch.empty(4 * tot_c, dtype=torch.float32, device=dev)   # scale, shift, mean, invstd
        off = 0
        for l in self.layers:
            c = l.cout
            l.stat_sum = self.acc64[off:off + c]
            l.stat_sumsq = self.acc64[tot_c + off:tot_c + off + c]
            l.scale = self.vec32[off:off + c]
            l.shift = self.vec32[tot_c + off:tot_c + off + c]
            l.mean = self.vec32[2 * tot_c + off:2 * tot_c + off + c]
            l.invstd = self.vec32[3 * tot_c + off:3 * tot_c + off + c]
            off += c
        self._tot_c = tot_c
        for l in self.layers:
            if l.first and self.pair_first:
                l.wf = torch.zeros((2 * l.cout, 2 * self.kpad), dtype=dt, device=dev)
            elif l.first:
                l.wf = torch.zeros((l.cout, self.kpad), dtype=dt, device=dev)
            else:
                l.wf = torch.empty((l.cout, 9, l.cin), dtype=dt, device=dev)
                l.wd = torch.empty((l.cin, 9, l.cout), dtype=dt, device=dev)
        for nd in self.nodes:
            if isinstance(nd, ConvT):
                nd.wf = torch.empty((4 * nd.cout, nd.cin), dtype=dt, device=dev)
                nd.wd = torch.empty((nd.cin, 4, nd.cout), dtype=dt, device=dev)
        return self

    # ---- weights ----------------------------------------------------------------------------------------------------
    def pack_weights(self):
        """fp32 OIHW / IOHW parameters -> K-major operand packs: one batched launch, only when a parameter changed."""
        params = list(self.model.parameters())
        versions = tuple(p._version for p in params)
        ptrs = tuple(p.data_ptr() for p in params)
        if (versions, ptrs) == self._pack_versions:
            return
        if self._pack_jobs is None or self._pack_ptrs != ptrs:
            jobs = []
            for nd in self.nodes:
                if isinstance(nd, ConvBNReLU):
                    w = nd.conv.weight
                    if nd.first and self.pair_first:
                        jobs.append((w.data_ptr(), nd.wf.data_ptr(), None, 3, nd.cout, nd.cin, 2 * self.kpad))
                    elif nd.first:
                        jobs.append((w.data_ptr(), nd.wf.data_ptr(), None, 2, nd.cout, nd.cin, self.kpad))
                    else:
                        jobs.append((w.data_ptr(), nd.wf.data_ptr(), nd.wd.data_ptr(), 0, nd.cout, nd.cin, 0))
                elif isinstance(nd, ConvT):
                    jobs.append((nd.mod.weight.data_ptr(), nd.wf.data_ptr(), nd.wd.data_ptr(), 1, nd.cout, nd.cin, 0))
            self._pack_jobs = L.WeightJobs(jobs, self.device) if jobs else None
            self._pack_ptrs = ptrs
        if self._pack_jobs is not None:
            L.weights_pack(self._pack_jobs, self.dt)
        for fn in self._extra_pack:
            fn()
        self._pack_versions = (versions, ptrs)

    # ---- forward ----------------------------------------------------------------------------------------------------
    def _bn_args(self, l: ConvBNReLU, training: bool):
        bn = l.bn
        if bn.momentum is None:
            # nn.BatchNorm2d(momentum=None) means a cumulative moving average whose factor 1/num_batches_tracked lives on
            # the device; reading it would stall the stream (and break graph capture), so it is refused loudly
            raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative average) is not supported by the fused engine; "
                                      "the reference always uses the default momentum 0.1 (unet/unet.py:17,20)")
        track = bn.track_running_stats and bn.running_mean is not None
        if bn.weight is None or bn.bias is None:
            raise NotImplementedError("BatchNorm2d(affine=False) is not supported by the fused engine")
        return (l.stat_sum, l.stat_sumsq, self.n * l.z.shape[1] * l.z.shape[2], l.cout, training, bn.weight, bn.bias,
                l.conv.bias, bn.running_mean if track else None, bn.running_var if track else None,
                bn.num_batches_tracked if (track and training) else None, bn.momentum, bn.eps, l.scale, l.shift, l.mean,
                l.invstd)

    @staticmethod
    def _pairs(t):
        """[N,H,W,C] contiguous -> the same memory as [N,H,W/2,2C] (two horizontally adjacent pixels per row)."""
        n, h, w, c = t.shape
        return t.view(n, h, w // 2, 2 * c)

    def _build_forward(self, training: bool):
        calls = []
        n = self.n
        for nd in self.nodes:
            if isinstance(nd, ConvBNReLU):
                l = nd
                count = n * l.z.shape[1] * l.z.shape[2]
                use_batch_stats = training or not (l.bn.track_running_stats and l.bn.running_mean is not None)
                flops = 2 * count * 9 * l.cin * l.cout
                if l.first and self.pair_first:
                    fs = self.first_stats
                    calls.append(L.prep_conv(self._pairs(l.src.t), l.wf, self._pairs(l.z), L.MODE_1X1,
                                             stat_sum=fs[:128] if use_batch_stats else None,
                                             stat_sumsq=fs[128:] if use_batch_stats else None, algo=self.algo, algo_flops=flops,
                                             label=l.name))
                    if use_batch_stats:     # the two pixels of a pair are the same 64 BatchNorm channels
                        calls.append(_TorchCall(lambda fs=fs, l=l: (torch.add(fs[:64], fs[64:128], out=l.stat_sum),
                                                                    torch.add(fs[128:192], fs[192:], out=l.stat_sumsq))))
                else:
                    calls.append(L.prep_conv(l.src.t, l.wf, l.z, L.MODE_1X1 if l.first else L.MODE_3X3,
                                             stat_sum=l.stat_sum if use_batch_stats else None,
                                             stat_sumsq=l.stat_sumsq if use_batch_stats else None, algo=self.algo,
                                             algo_flops=flops if l.first else None, label=l.name))
                calls.append(L.prep_bn_finalize(*self._bn_args(l, use_batch_stats), label=l.name))
                if l.fused_head is not None:
                    hd = l.fused_head
                    calls.append(L.prep_bn_relu_head_fprop(l.z, l.scale, l.shift, None, hd.conv.weight, hd.conv.bias, hd.dout,
                                                           self.output, label=l.name))
                else:
                    calls.append(L.prep_bn_relu_apply(l.z, l.scale, l.shift, l.out.t, l.pooled.t if l.pooled else None,
                                                      l.pool_idx, label=l.name))
            elif isinstance(nd, ConvT):
                calls.append(L.prep_conv(nd.src.t, nd.wf, nd.out.t, L.MODE_CONVT, bias=nd.mod.bias, algo=self.algo, label=nd.name))
            elif isinstance(nd, Head):
                if not nd.fused:
                    calls.append(L.prep_head_fprop(nd.src.t, nd.conv.weight, nd.conv.bias, nd.dout, self.output, label=nd.name))
            else:
                calls += nd.forward_calls(self, training)
        return calls

    def forward(self, inputs: Sequence[torch.Tensor], training: bool) -> torch.Tensor:
        self.generation += 1
        self.training_pass = training
        self.pack_weights()
        key = bool(training)
        if key not in self._fwd_calls:
            self._fwd_calls[key] = self._build_forward(training)
        self.acc64.zero_()
        stream = L.stream_ptr()
        for fn in self.input_call_builders:
            fn(inputs, stream)
        for c in self._fwd_calls[key]:
            c(stream)
        return self.output

    # ---- backward ---------------------------------------------------------------------------------------------------
    def _plan_backward(self):
        """Gradient buffers, fusion decisions and the prepared backward launch list (built on the first backward)."""
        dev = self.device
        m = self.model
        # 1. where gradients live -------------------------------------------------------------------------------------
        for cat in self.cats:
            if cat.needs_grad and cat.grad is None:
                if len(cat.consumers) != 1:
                    raise NotImplementedError(f"{cat.name}: a concat buffer needs exactly one consumer")
                cat.grad = torch.empty_like(cat.t)
                for p in cat.parts:
                    p.grad = cat.grad[..., p.offset:p.offset + p.c]
        for a in self.feature_inputs:
            if a.needs_grad and a.parent is None and a.grad is None and a.consumers:
                a.grad = torch.empty_like(a.t)
        for nd in self.nodes:
            if isinstance(nd, ConvBNReLU):
                outs = [nd.out] + ([nd.pooled] if nd.pooled is not None else [])
            elif isinstance(nd, ConvT):
                outs = [nd.out]
            elif isinstance(nd, Head):
                outs = []
            else:
                outs = nd.outputs()
            for a in outs:
                if len(a.consumers) > 1 or (a.consumers and a.parent is not None):
                    raise NotImplementedError(f"{a.name}: an activation with several consumers is not supported")
                if a.parent is not None or not a.needs_grad or a.grad is not None or not a.consumers:
                    continue
                c0 = a.consumers[0]
                if isinstance(c0, Head) and c0.fused:
                    continue            # the fused head backward never materialises d(activation)
                a.grad = torch.empty_like(a.t)
        # 2. BatchNorm-backward reduction fused into the launch that produces d(out) -------------------------------------
        for nd in self.nodes:
            if not nd.can_fuse_reduce or not self._node_runs_backward(nd):
                continue
            src = nd.src
            p = src.producer
            if (self.fuse_reduce_enabled and isinstance(p, ConvBNReLU) and src is p.out and src.parent is None and src.needs_grad
                    and (p.pooled is None or not (p.pooled.needs_grad and p.pooled.consumers))):
                nd.fuse_reduce_of = p
                p.reduced = True
        # 3. flat gradient layout in backward completion order -----------------------------------------------------------
        self.grad_order: List[torch.nn.Parameter] = []
        order_nodes = [nd for nd in reversed(self.nodes)]
        for nd in order_nodes:
            if self._node_runs_backward(nd):
                self.grad_order += [p for p in nd.params() if p.requires_grad]
        sizes = [p.numel() for p in self.grad_order]
        self.grad_offsets = [0]
        for s in sizes:
            self.grad_offsets.append(self.grad_offsets[-1] + s)
        self.grad_total = self.grad_offsets[-1]
        self._off_of = {id(p): o for p, o in zip(self.grad_order, self.grad_offsets[:-1])}
        # 4. weight-gradient workspaces (operand layout, fp32) ------------------------------------------------------------
        ws_sizes, ws_nodes = [], []
        for nd in order_nodes:
            if not self._node_runs_backward(nd):
                continue
            if isinstance(nd, ConvBNReLU) and nd.conv.weight.requires_grad:
                if nd.first and self.pair_first:
                    ws_sizes.append(2 * nd.cout * 2 * self.kpad)
                else:
                    ws_sizes.append(nd.cout * (self.kpad if nd.first else 9 * nd.cin))
                ws_nodes.append(nd)
            elif isinstance(nd, ConvT) and nd.mod.weight.requires_grad:
                ws_sizes.append(nd.cin * 4 * nd.cout)
                ws_nodes.append(nd)
            elif hasattr(nd, "ws_size") and nd.ws_size():
                ws_sizes.append(nd.ws_size())
                ws_nodes.append(nd)
        self.ws = torch.empty(max(1, sum(ws_sizes)), dtype=torch.float32, device=dev)
        off = 0
        for nd, s in zip(ws_nodes, ws_sizes):
            nd.ws = self.ws[off:off + s]
            off += s
        # backward reduction sums [2][C] per layer (+ the fused head's (3 + dout) rows)
        head_rows = (3 + self.head.dout) if (self.head is not None and self.head.fused) else 0
        head_c = self.head.src.c if head_rows else 0
        self.bwd64 = torch.zeros(2 * self._tot_c + head_rows * head_c, dtype=torch.float64, device=dev)
        self.head_sums = self.bwd64[2 * self._tot_c:]
        off = 0
        for l in self.layers:
            l.bwd_sums = self.bwd64[off:off + 2 * l.cout]
            off += 2 * l.cout
        for l in self.layers:
            if self._node_runs_backward(l):
                l.dz = torch.empty_like(l.z)
        # 5. the launch list ---------------------------------------------------------------------------------------------
        self._det_sites = []
        self._bwd = self._build_backward_calls(order_nodes)
        self._attach_partial_scratch()

    def _node_runs_backward(self, nd) -> bool:
        if isinstance(nd, ConvBNReLU):
            return nd.out.needs_grad
        if isinstance(nd, ConvT):
            return nd.out.needs_grad
        if isinstance(nd, Head):
            return any(p.requires_grad for p in nd.params()) or nd.src.needs_grad
        return nd.runs_backward()

    def _frozen_bn_scratch(self, c: int):
        """(dgamma, dbeta) device pointers for a BatchNorm whose affine parameters are frozen while gradients still flow
        through it: the kernels always write both vectors, so they get a throw-away buffer."""
        buf = getattr(self, "_scratch_c", None)
        if buf is None or buf.numel() < 2 * c:
            buf = self._scratch_c = torch.empty(max(8192, 2 * c), dtype=torch.float32, device=self.device)
        return buf.data_ptr(), buf.data_ptr() + 4 * c

    def _patch_bn_grads(self, call, args, l):
        """dgamma / dbeta of layer l land in the flat gradient buffer when trainable, in a throw-away buffer when frozen."""
        args.dgamma, args.dbeta = self._frozen_bn_scratch(l.cout)
        if l.bn.weight.requires_grad:
            call.patch_ptr(args, "dgamma", self._goff(l.bn.weight))
        if l.bn.bias.requires_grad:
            call.patch_ptr(args, "dbeta", self._goff(l.bn.bias))

    def _goff(self, p) -> int:
        """Byte offset of parameter p's gradient in the flat buffer."""
        return 4 * self._off_of[id(p)]

    def _bn_red(self, p: Optional[ConvBNReLU]):
        return (p.z, p.scale, p.shift, p.mean, p.invstd, p.bwd_sums) if p is not None else None

    def wgrad_call(self, u, s, dw, mode, **kw) -> L.Call:
        """Prepared unetk_wgrad launch; in deterministic mode the call site is recorded so that one shared partial-sum
        scratch buffer (sized for the largest site; the launches run one after another) can be attached afterwards."""
        call = L.prep_wgrad(u, s, dw, mode, algo=self.algo, **kw)
        if self.deterministic and u.dtype == torch.bfloat16 and self.algo in (L.ALGO_AUTO, L.ALGO_TC):
            self._det_sites.append((call.keep[0], L.wgrad_partial_bytes(u, s, mode, self.algo)))
        return call

    def channel_sum_call(self, t, label) -> L.Call:
        """Prepared per-channel sum (bias gradients) whose destination is patched per backward pass (argument 1)."""
        scratch = None
        if self.deterministic:
            if getattr(self, "_csum_scratch", None) is None or self._csum_scratch.numel() < L.CHANNEL_SUM_BLOCKS * t.shape[3]:
                self._csum_scratch = torch.empty(L.CHANNEL_SUM_BLOCKS * max(t.shape[3], 1024), dtype=torch.float32, device=self.device)
            scratch = self._csum_scratch
        return L.prep_channel_sum(t, None, label=label, scratch=scratch)

    def _attach_partial_scratch(self):
        need = max([n for _, n in self._det_sites], default=0)
        if need == 0:
            return
        self._partial = torch.empty(need // 4, dtype=torch.float32, device=self.device)
        for args, n in self._det_sites:
            if n > 0:
                args.partial = self._partial.data_ptr()
                args.partial_bytes = need
                args.algo |= L.TC_DETERMINISTIC

    def _dgrad_call(self, nd, mode, dy: torch.Tensor, label):
        """Data gradient of a conv3x3 / convT node into its source's gradient view; a source that is a concat of
        [trainable, frozen] parts gets only the channel range that needs it."""
        src = nd.src
        if not src.needs_grad:
            return None
        if isinstance(src, Cat):
            need = [p for p in src.parts if p.needs_grad]
            lo, hi = need[0].offset, need[-1].offset + need[-1].c
            contiguous = all(a.offset + a.c == b.offset for a, b in zip(need, need[1:]))
            if contiguous and (lo > 0 or hi < src.c) and mode == L.MODE_3X3:
                wd = nd.wd[lo:hi]                      # rows of the [ci][8-t][co] pack = input channels
                return L.prep_conv(dy, wd, src.grad[..., lo:hi], mode, algo=self.algo, label=label)
        return L.prep_conv(dy, nd.wd, src.grad, mode, algo=self.algo, bn_reduce=self._bn_red(nd.fuse_reduce_of), label=label)

    def _build_backward_calls(self, order_nodes):
        """List of (callable(stream, flat_ptr)) in backward order; bucket boundaries are ('bucket', seg_index, end_offset)."""
        ops: list = []
        jobs_pending: list = []
        self._unpack_jobs: list = []
        all_jobs: list = []
        done_params = 0
        closed_at = 0

        def close_segment():
            nonlocal closed_at
            seg = len(self._unpack_jobs)
            self._unpack_jobs.append(L.WeightJobs(list(jobs_pending), self.device) if jobs_pending else None)
            ops.append(("bucket", seg, self.grad_offsets[done_params]))
            all_jobs.extend(jobs_pending)
            jobs_pending.clear()
            closed_at = done_params

        for nd in order_nodes:
            if not self._node_runs_backward(nd):
                continue
            if isinstance(nd, Head):
                hd = nd
                if not hd.fused:
                    wreq = hd.conv.weight.requires_grad
                    if not wreq:
                        raise NotImplementedError("a frozen classifier head is not supported")
                    hd.g_in = hd.src.grad
                    call, box = L.prep_head_bwd(self._dlogits_slot, hd.src.t, hd.conv.weight, hd.dout, hd.g_in, label=hd.name)
                    call.patch_ptr(box, "dw", self._goff(hd.conv.weight))
                    if hd.conv.bias is not None and hd.conv.bias.requires_grad:
                        call.patch_ptr(box, "db", self._goff(hd.conv.bias))
                    ops.append(call)
                    done_params += len([p for p in hd.params() if p.requires_grad])
                    ops.append(("bucket_only", None, self.grad_offsets[done_params]))
                else:
                    done_params += len([p for p in hd.params() if p.requires_grad])
                continue
            if isinstance(nd, ConvBNReLU):
                l = nd
                if l.fused_head is not None:
                    hd = l.fused_head
                    args, calls = L.prep_head_bn_bwd(self._dlogits_slot, l.z, hd.conv.weight, hd.dout, l.scale, l.shift, l.mean,
                                                     l.invstd, self.head_sums, l.dz, label=l.name)
                    ap = calls[1]
                    self._patch_bn_grads(ap, args, l)
                    if hd.conv.weight.requires_grad:
                        ap.patch_ptr(args, "dw_head", self._goff(hd.conv.weight))
                    else:       # frozen classifier: the kernel still needs somewhere to put dW (dout x C floats)
                        self._head_scratch = torch.empty(hd.dout * l.cout, dtype=torch.float32, device=self.device)
                        args.dw_head = self._head_scratch.data_ptr()
                    if hd.conv.bias is not None and hd.conv.bias.requires_grad:
                        ap.patch_ptr(args, "db_head", self._goff(hd.conv.bias))
                    ops += calls
                    ops.append(("bucket_only", None, self.grad_offsets[done_params]))
                else:
                    dy = l.out.grad
                    dpool = l.pooled.grad if (l.pooled is not None and l.pooled.grad is not None) else None
                    if dy is None and dpool is None:
                        raise RuntimeError(f"{l.name}: no gradient reaches this layer")
                    args, calls = L.prep_bn_relu_bwd(l.z, dy, dpool, l.scale, l.shift, l.mean, l.invstd, l.bwd_sums, l.dz,
                                                     pool_idx=l.pool_idx if dpool is not None else None,
                                                     reduced=l.reduced and dpool is None, label=l.name)
                    self._patch_bn_grads(calls[-1], args, l)
                    ops += calls
                wreq = l.conv.weight.requires_grad
                flops = 2 * self.n * l.z.shape[1] * l.z.shape[2] * 9 * l.cin * l.cout
                if l.first:
                    if wreq:
                        if self.pair_first:
                            # dW' [2*cout][2*kpad] over pixel pairs; weights_unpack (kind 3) adds its two diagonal blocks
                            ops.append(self.wgrad_call(self._pairs(l.dz), self._pairs(l.src.t), l.ws, 0, algo_flops=flops, label=l.name))
                            jobs_pending.append((l.ws.data_ptr(), self._goff(l.conv.weight), None, 3, l.cout, l.cin, 2 * self.kpad))
                        else:
                            ops.append(self.wgrad_call(l.dz, l.src.t, l.ws, 0, algo_flops=flops, label=l.name))
                            jobs_pending.append((l.ws.data_ptr(), self._goff(l.conv.weight), None, 2, l.cout, l.cin, self.kpad))
                else:
                    dg = self._dgrad_call(l, L.MODE_3X3, l.dz, l.name)
                    if dg is not None:
                        ops.append(dg)
                    if wreq:
                        ops.append(self.wgrad_call(l.dz, l.src.t, l.ws, 1, label=l.name))
                        jobs_pending.append((l.ws.data_ptr(), self._goff(l.conv.weight), None, 0, l.cout, l.cin, 0))
                # conv bias in front of train-mode BatchNorm: its gradient is identically zero (flat buffer is zeroed)
                done_params += len([p for p in l.params() if p.requires_grad])
            elif isinstance(nd, ConvT):
                ct = nd
                g_out = ct.out.grad
                dg = self._dgrad_call(ct, L.MODE_CONVT_GATHER, g_out, ct.name)
                if dg is not None:
                    ops.append(dg)
                if ct.mod.weight.requires_grad:
                    ops.append(self.wgrad_call(ct.src.t, g_out, ct.ws, 2, label=ct.name))
                    jobs_pending.append((ct.ws.data_ptr(), self._goff(ct.mod.weight), None, 1, ct.cout, ct.cin, 0))
                if ct.mod.bias is not None and ct.mod.bias.requires_grad:
                    ops.append(_PatchedArg(self.channel_sum_call(g_out, ct.name), 1, self._goff(ct.mod.bias)))
                done_params += len([p for p in ct.params() if p.requires_grad])
            else:
                done_params += nd.backward_calls(self, ops, jobs_pending)
            if nd.end_block:
                close_segment()
        if jobs_pending or closed_at != done_params or not self._unpack_jobs:
            close_segment()
        self._unpack_all = L.WeightJobs(all_jobs, self.device) if all_jobs else None
        return ops

    def backward(self, dlogits: torch.Tensor, bucket_hook: Optional[Callable] = None) -> Dict[torch.nn.Parameter, torch.Tensor]:
        """Returns {parameter: gradient}; the gradients are views of one freshly allocated flat buffer."""
        if not self.training_pass and any(l.bn.track_running_stats for l in self.layers):
            raise RuntimeError("backward through an eval-mode (running statistics) forward is not supported; "
                               "call model.train() before the forward pass")
        if self._bwd is None:
            # the upstream gradient is copied into a plan-owned slot so that prepared calls keep a stable pointer
            self._dlogits_slot = torch.empty_like(self.output)
            self._plan_backward()
        flat = torch.zeros(max(1, self.grad_total), dtype=torch.float32, device=self.device)
        g = {}
        for p, o0, o1 in zip(self.grad_order, self.grad_offsets[:-1], self.grad_offsets[1:]):
            g[p] = flat[o0:o1].view(p.shape)
        self.ws.zero_()
        self.bwd64.zero_()
        self.flat_grad = flat
        self._dlogits_slot.copy_(dlogits)
        stream = L.stream_ptr()
        base = flat.data_ptr()
        # block-by-block un-pack + hook only when a data-parallel exchange wants early buckets; otherwise ONE batched
        # un-pack after the last weight gradient (the per-block launches are latency-bound, 2-3 blocks per SM each)
        per_segment = bucket_hook is not None and getattr(self.model, "_bucket_per_segment", True)
        for op in self._bwd:
            if isinstance(op, tuple):
                kind, seg, end = op
                if per_segment:
                    if kind == "bucket" and self._unpack_jobs[seg] is not None:
                        L.weights_unpack(self._unpack_jobs[seg], flat)
                    bucket_hook(self, end)
                continue
            op(stream, base)
        if not per_segment and self._unpack_all is not None:
            L.weights_unpack(self._unpack_all, flat)
        # gradients of NCHW feature-map inputs (block-level API)
        self.input_grads = {}
        for a in self.feature_inputs:
            if a.needs_grad and a.grad is not None:
                n, h, w, c = a.t.shape
                gi = torch.empty((n, c, h, w), dtype=torch.float32, device=self.device)
                L.prep_nhwc_to_nchw(a.grad, gi)(stream)
                self.input_grads[a.input_index] = gi
        return g


class _TorchCall:
    """A tiny torch-side step inside a prepared launch list (runs on the current stream, captured by CUDA graphs)."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, stream, base_ptr=0):
        self.fn()


class _PatchedArg:
    """A prepared call whose positional argument `index` is a pointer into the per-backward flat gradient buffer."""

    def __init__(self, call: L.Call, index: int, byte_offset: int):
        self.call, self.index, self.off = call, index, byte_offset

    def __call__(self, stream, base_ptr=0):
        a = list(self.call.args)
        a[self.index] = base_ptr + self.off
        self.call.args = tuple(a)
        self.call(stream)


# =====================================================================================================================
# autograd node + plan cache
# =====================================================================================================================
class _NetFunction(torch.autograd.Function):
    """The single autograd node standing for a whole network forward (e.g. unet.forward, unet/unet.py:93-105)."""

    @staticmethod
    def forward(ctx, model, plan, n_inputs, *tensors):
        ctx.plan, ctx.model = plan, model
        ctx.n_inputs = n_inputs
        ctx.params = tensors[n_inputs:]
        ctx.generation = plan.generation + 1
        # the result leaves the plan as a COPY: the plan-owned buffer is overwritten by the next forward pass (and returning
        # the same tensor object every step would keep the previous step's autograd graph alive, which breaks CUDA-graph
        # capture); 4 bytes x N x C x H x W once per step
        return plan.forward(tensors[:n_inputs], training=model.training).clone()

    @staticmethod
    def backward(ctx, dout):
        plan: NetPlan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("the activations of this forward pass were overwritten by a later forward pass")
        hook = getattr(ctx.model, "_bucket_hook", None)
        scale = getattr(ctx.model, "_grad_scale", None)
        if scale is not None and scale != 1.0:
            dout = dout * scale
        grads = plan.backward(dout, hook)
        plan.busy = False
        done = getattr(ctx.model, "_backward_done_hook", None)
        if done:
            done(plan)
        out = [grads.get(p) if p.requires_grad else None for p in ctx.params]
        gin = [plan.input_grads.get(i) if ctx.needs_input_grad[3 + i] else None for i in range(ctx.n_inputs)]
        return (None, None, None, *gin, *out)


class EngineModule(torch.nn.Module):
    """Base of every engine-backed module: one launch-plan engine per instance, dropped whenever parameters move or are
    re-typed, never pickled."""
    _engine = None

    def _apply(self, fn, *args, **kwargs):
        # .to()/.cuda()/.double() move or retype parameters: cached device buffers are then stale
        self._engine = None
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        return state


class Engine:
    """Plan cache + entry point used by a model's ``forward``.  ``build(plan, *inputs)`` describes the network."""

    def __init__(self, model, build: Callable, check_inputs: Optional[Callable] = None):
        self.model, self.build, self.check_inputs = model, build, check_inputs
        self.plans: Dict[tuple, List[NetPlan]] = {}

    def run(self, *inputs: torch.Tensor) -> torch.Tensor:
        m = self.model
        params = list(m.parameters())
        L.require_cuda(*inputs, params[0])
        for x in inputs:
            if x.device != params[0].device:
                raise RuntimeError(f"input is on {x.device} but the model is on {params[0].device}")
        for p in params:
            if p.dtype != torch.float32:
                raise RuntimeError("parameters must stay fp32 (master weights); precision is selected with model.precision")
        if self.check_inputs is not None:
            self.check_inputs(*inputs)
        precision = os.environ.get("UNETK_PRECISION", getattr(m, "precision", "bf16"))
        inputs = tuple(x.contiguous() if x.dtype == torch.float32 else x.float().contiguous() for x in inputs)
        dev = inputs[0].device
        key = (tuple(tuple(x.shape) for x in inputs), tuple(bool(x.requires_grad) for x in inputs), precision, dev.index)
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or any(x.requires_grad for x in inputs))
        pool = self.plans.setdefault(key, [])
        plan = next((p for p in pool if not p.busy), None)
        if plan is None and len(pool) >= 2:
            # forwards whose backward never ran (e.g. a validation loss computed with grad enabled) must not leak
            # buffers: recycle the oldest plan; a late backward on it fails the generation check loudly
            plan = min(pool, key=lambda p: p.generation)
        if plan is None:
            with torch.cuda.device(dev):
                plan = NetPlan(m, inputs[0].shape[0], precision, dev)
                self.build(plan, *inputs)
                plan.finish()
            pool.append(plan)
        algo = {"auto": L.ALGO_AUTO, "simt": L.ALGO_SIMT, "tc": L.ALGO_TC}[os.environ.get("UNETK_ALGO", getattr(m, "conv_algo", "auto"))]
        if algo != plan.algo:
            # the kernel selection is baked into the prepared calls: rebuild them (debug / cross-check knob, rare)
            plan.algo = algo
            plan._fwd_calls.clear()
            plan._bwd = None
        with torch.cuda.device(dev):
            if needs_grad:
                plan.busy = True
                return _NetFunction.apply(m, plan, len(inputs), *inputs, *params)
            with torch.no_grad():
                return plan.forward(inputs, training=m.training).clone()
