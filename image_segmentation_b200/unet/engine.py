"""Whole-network forward/backward of the U-Net over libunetk.so (NHWC, bf16 or fp32 activations).

The reference runs ``unet.forward`` (unet/unet.py:93-105) as ~90 ATen calls and lets autograd replay
~150 more.  Here one ``torch.autograd.Function`` owns the complete pass: it packs the fp32
``nn.Parameter`` tensors into K-major operand layouts, walks the 23 contraction layers through the
C ABI (conv3x3 / ConvTranspose2d / 1x1 head), keeps BatchNorm + ReLU + MaxPool as fused bandwidth
kernels around them, writes skip connections and up-sampled maps straight into one concat buffer
(the ``torch.cat`` at unet/unet.py:63 becomes a channel-slice view), and returns parameter gradients
in PyTorch's own layouts so that ``optimizer.step()`` / gradient accumulation (utils/training.py:49-56)
work unchanged.

Host code is plumbing: PyTorch allocates, this file sequences launches on the current stream.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import torch

from .. import _lib as L

ENC_CH = (64, 128, 256, 512, 1024)
BN_EPS_DEFAULT = 1e-5


def _dtype_of(precision: str):
    if precision == "bf16":
        return torch.bfloat16
    if precision == "fp32":
        return torch.float32
    raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")


class _ConvBN:
    """One conv3x3 -> BatchNorm -> ReLU layer: parameter holders, operand packs and activations."""

    def __init__(self, name, conv, bn, cin, cout, h, w, first):
        self.name, self.conv, self.bn = name, conv, bn
        self.cin, self.cout, self.h, self.w, self.first = cin, cout, h, w, first
        self.kin = cin  # channels of the gathered operand (im2col width for the first layer)
        self.src = None      # input activation view [N,h,w,kin]
        self.z = None        # raw conv output
        self.a = None        # activated output view
        self.pooled = None   # optional max-pooled output
        self.pool_idx = None  # 2-bit arg-max position per pooled element (written forward, read backward)
        self.wf = self.wd = None
        self.dz = None       # gradient wrt z
        self.g_in = None     # where the data gradient (wrt src) is written; None for the first layer


class _ConvT:
    def __init__(self, name, mod, cin, cout, h, w):
        self.name, self.mod, self.cin, self.cout, self.h, self.w = name, mod, cin, cout, h, w
        self.src = None   # [N,h,w,cin]
        self.out = None   # view [N,2h,2w,cout] inside the concat buffer
        self.wf = self.wd = None
        self.g_out = None  # gradient wrt out (view of dcat)
        self.g_in = None   # gradient wrt src


class UNetPlan:
    """Device buffers + launch sequence for one (N, H, W, din, dout, precision) problem."""

    def __init__(self, model, n, h, w, precision, device):
        if h % 16 or w % 16:
            # the reference fails in torch.cat (unet/unet.py:63) for such sizes
            raise ValueError(f"U-Net input height/width must be multiples of 16, got {h}x{w}")
        self.model, self.n, self.h, self.w = model, n, h, w
        self.precision, self.device = precision, device
        self.dt = _dtype_of(precision)
        self.din, self.dout = model.din, model.dout
        self.busy = False
        self.generation = 0
        self.algo = L.ALGO_AUTO
        self._pack_versions = None
        self._pack_jobs = None
        self._pack_ptrs = None
        self._unpack_jobs = None
        self._build()

    # ------------------------------------------------------------------------------------------
    def _act(self, n, h, w, c):
        return torch.empty((n, h, w, c), dtype=self.dt, device=self.device)

    @staticmethod
    def _pairs(t):
        """[N,H,W,C] contiguous -> the same memory as [N,H,W/2,2C] (two horizontally adjacent pixels per row)."""
        n, h, w, c = t.shape
        return t.view(n, h, w // 2, 2 * c)

    def _build(self):
        m, n, dev, dt = self.model, self.n, self.device, self.dt
        H, W = self.h, self.w
        self.kpad = ((9 * self.din + 63) // 64) * 64
        # First layer in pixel-pair form (bf16 tier): with K padded to 32 per pixel, two horizontally adjacent pixels make
        # one 64-wide GEMM row, and the block-diagonal weight [[W 0] [0 W]] (N = 128) writes both pixels' 64 output
        # channels -- which IS the NHWC layout of the pair.  The im2col buffer and its two readers (this conv and the
        # first-layer weight gradient) move half the bytes of the K = 64 padding; all three kernels are HBM-bound.
        self.pair_first = (dt == torch.bfloat16 and W % 2 == 0 and 9 * self.din <= 32
                           and os.environ.get("UNETK_FIRST_PAIR", "1") == "1")
        if self.pair_first:
            self.kpad = 32
        self.xcol = self._act(n, H, W, self.kpad)
        hs = [H >> i for i in range(5)]
        ws = [W >> i for i in range(5)]
        self.cat = [self._act(n, hs[l], ws[l], 2 * ENC_CH[l]) for l in range(4)]
        self.dcat = [None] * 4
        self.layers: List[_ConvBN] = []
        self.convts: List[_ConvT] = []

        def dc_modules(block):
            seq = block.doubleConvReLU
            return (seq[0], seq[1]), (seq[3], seq[4])

        # ---- encoder ----
        enc_blocks = [m.down1, m.down2.maxpool_doubleConv[1], m.down3.maxpool_doubleConv[1],
                      m.down4.maxpool_doubleConv[1], m.down5.maxpool_doubleConv[1]]
        enc_names = ["down1", "down2", "down3", "down4", "down5"]
        prev = self.xcol
        self.enc: List[List[_ConvBN]] = []
        for l in range(5):
            c = ENC_CH[l]
            cin = self.din if l == 0 else ENC_CH[l - 1]
            (c1, b1), (c2, b2) = dc_modules(enc_blocks[l])
            l1 = _ConvBN(enc_names[l] + ".c1", c1, b1, cin, c, hs[l], ws[l], first=(l == 0))
            if l == 0:
                l1.kin = self.kpad
            l1.src = prev
            l1.z = self._act(n, hs[l], ws[l], c)
            l1.a = self._act(n, hs[l], ws[l], c)
            l2 = _ConvBN(enc_names[l] + ".c2", c2, b2, c, c, hs[l], ws[l], first=False)
            l2.src = l1.a
            l2.z = self._act(n, hs[l], ws[l], c)
            if l < 4:
                l2.a = self.cat[l][..., :c]
                l2.pooled = self._act(n, hs[l + 1], ws[l + 1], c)
                l2.pool_idx = torch.empty((n, hs[l + 1], ws[l + 1], c // 8), dtype=torch.int16, device=dev)
                prev = l2.pooled
            else:
                l2.a = self._act(n, hs[l], ws[l], c)
                prev = l2.a
            self.enc.append([l1, l2])
            self.layers += [l1, l2]
        # ---- decoder (up1 works at level index 3, up4 at level index 0) ----
        ups = [m.up1, m.up2, m.up3, m.up4]
        self.dec: List[List[_ConvBN]] = []
        for i, up in enumerate(ups):
            l = 3 - i
            c = ENC_CH[l]
            ct = _ConvT(f"up{i + 1}.upsample", up.upsample, 2 * c, c, hs[l + 1], ws[l + 1])
            ct.src = prev
            ct.out = self.cat[l][..., c:]
            self.convts.append(ct)
            (c1, b1), (c2, b2) = dc_modules(up.doubleConv)
            l1 = _ConvBN(f"up{i + 1}.c1", c1, b1, 2 * c, c, hs[l], ws[l], first=False)
            l1.src = self.cat[l]
            l1.z = self._act(n, hs[l], ws[l], c)
            l1.a = self._act(n, hs[l], ws[l], c)
            l2 = _ConvBN(f"up{i + 1}.c2", c2, b2, c, c, hs[l], ws[l], first=False)
            l2.src = l1.a
            l2.z = self._act(n, hs[l], ws[l], c)
            l2.a = self._act(n, hs[l], ws[l], c)
            prev = l2.a
            self.dec.append([l1, l2])
            self.layers += [l1, l2]
        self.head_in = prev

        # ---- per-channel scratch: fp64 accumulators (zeroed once per pass) and fp32 vectors ----
        tot_c = sum(l.cout for l in self.layers)
        # sum, sumsq, bwd s1, s2 (+ 2 x 128 for the pixel-pair first layer, whose statistics arrive as two halves)
        self.acc64 = torch.zeros(4 * tot_c + 256, dtype=torch.float64, device=dev)
        self.first_stats = self.acc64[4 * tot_c:]
        self.vec32 = torch.empty(4 * tot_c, dtype=torch.float32, device=dev)  # scale, shift, mean, invstd
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
        # backward reduction sums [2][C] per layer, contiguous per layer
        # + the sums of the fused head/BatchNorm backward: (3 + dout) rows of the last block's channel count
        self.head_rows = 3 + self.dout
        self.bwd64 = torch.zeros(2 * tot_c + self.head_rows * ENC_CH[0], dtype=torch.float64, device=dev)
        self.head_sums = self.bwd64[2 * tot_c:]
        off = 0
        for l in self.layers:
            l.bwd_sums = self.bwd64[off:off + 2 * l.cout]
            off += 2 * l.cout

        # ---- operand packs ----
        for l in self.layers:
            if l.first and self.pair_first:
                l.wf = torch.zeros((2 * l.cout, 2 * self.kpad), dtype=dt, device=dev)
            elif l.first:
                l.wf = torch.zeros((l.cout, self.kpad), dtype=dt, device=dev)
            else:
                l.wf = torch.empty((l.cout, 9, l.cin), dtype=dt, device=dev)
                l.wd = torch.empty((l.cin, 9, l.cout), dtype=dt, device=dev)
        for ct in self.convts:
            ct.wf = torch.empty((4 * ct.cout, ct.cin), dtype=dt, device=dev)
            ct.wd = torch.empty((ct.cin, 4, ct.cout), dtype=dt, device=dev)

        self._bwd_ready = False
        # Optional (UNETK_WGRAD_STREAM=1): weight gradients on a side stream so that tensor-bound wgrad kernels may overlap
        # the HBM-bound BatchNorm backward of the next layer.  Measured on B200: 26.76 vs 26.83 ms/step -- the persistent
        # tcgen05 kernels own the shared memory of every SM, so almost nothing co-runs; off by default.
        self.side_stream = torch.cuda.Stream(device=dev) if os.environ.get("UNETK_WGRAD_STREAM", "0") == "1" else None

    # ------------------------------------------------------------------------------------------
    def _build_backward(self):
        """Gradient buffers are only allocated when a backward pass is actually requested."""
        n = self.n
        for l in range(4):
            c = ENC_CH[l]
            self.dcat[l] = self._act(n, self.h >> l, self.w >> l, 2 * c)
        for l in self.layers:
            l.dz = torch.empty_like(l.z)
        # data-gradient destinations
        for lvl in range(5):
            l1, l2 = self.enc[lvl]
            l2.g_in = torch.empty_like(l1.a)          # grad wrt l1.a
            l1.g_in = None if lvl == 0 else torch.empty_like(self.enc[lvl - 1][1].pooled)  # grad wrt pooled input
        for i in range(4):
            lvl = 3 - i
            l1, l2 = self.dec[i]
            l2.g_in = torch.empty_like(l1.a)
            l1.g_in = self.dcat[lvl]
            ct = self.convts[i]
            ct.g_out = self.dcat[lvl][..., ENC_CH[lvl]:]
            ct.g_in = torch.empty_like(ct.src)
        self.g_head_in = torch.empty_like(self.head_in)
        # flat layout of parameter gradients / weight-gradient workspaces in backward completion order
        self.grad_order: List[torch.nn.Parameter] = []
        m = self.model
        self.grad_order += [m.output.weight, m.output.bias]
        seq = []
        for i in (3, 2, 1, 0):                       # up4 ... up1
            l1, l2 = self.dec[i]
            seq += [l2, l1, self.convts[i]]
        for lvl in (4, 3, 2, 1, 0):
            l1, l2 = self.enc[lvl]
            seq += [l2, l1]
        self.bwd_seq = seq
        for item in seq:
            if isinstance(item, _ConvBN):
                self.grad_order += [item.conv.weight, item.conv.bias, item.bn.weight, item.bn.bias]
            else:
                self.grad_order += [item.mod.weight, item.mod.bias]
        sizes = [p.numel() for p in self.grad_order]
        self.grad_offsets = [0]
        for s in sizes:
            self.grad_offsets.append(self.grad_offsets[-1] + s)
        self.grad_total = self.grad_offsets[-1]
        # fp32 weight-gradient workspaces in operand layout ([Cu][taps][Cs])
        ws_sizes = []
        for item in seq:
            if isinstance(item, _ConvBN):
                if item.first and self.pair_first:
                    ws_sizes.append(2 * item.cout * 2 * self.kpad)
                else:
                    ws_sizes.append(item.cout * (self.kpad if item.first else 9 * item.cin))
            else:
                ws_sizes.append(item.cin * 4 * item.cout)
        self.ws_total = sum(ws_sizes)
        self.ws = torch.empty(self.ws_total, dtype=torch.float32, device=self.device)
        off = 0
        for item, s in zip(seq, ws_sizes):
            item.ws = self.ws[off:off + s]
            off += s
        # batched weight-gradient unpack tables, one per all-reduce segment (decoder level / encoder level);
        # destinations are byte offsets into the flat gradient buffer, which is a fresh allocation every backward pass
        off_of = {id(p): o for p, o in zip(self.grad_order, self.grad_offsets[:-1])}
        segments = []
        for i in (3, 2, 1, 0):
            segments.append([self.dec[i][1], self.dec[i][0], self.convts[i]])
        for lvl in (4, 3, 2, 1, 0):
            segments.append([self.enc[lvl][1], self.enc[lvl][0]])
        self._unpack_jobs = []
        all_jobs = []
        for items in segments:
            jobs = []
            for item in items:
                if isinstance(item, _ConvBN):
                    dst = 4 * off_of[id(item.conv.weight)]
                    if item.first and self.pair_first:
                        jobs.append((item.ws.data_ptr(), dst, None, 3, item.cout, item.cin, 2 * self.kpad))
                    elif item.first:
                        jobs.append((item.ws.data_ptr(), dst, None, 2, item.cout, item.cin, self.kpad))
                    else:
                        jobs.append((item.ws.data_ptr(), dst, None, 0, item.cout, item.cin, 0))
                else:
                    jobs.append((item.ws.data_ptr(), 4 * off_of[id(item.mod.weight)], None, 1, item.cout, item.cin, 0))
            self._unpack_jobs.append(L.WeightJobs(jobs, self.device))
            all_jobs += jobs
        # single process: one launch over every segment at the end of backward (the nine per-segment launches are
        # latency-bound, 2-3 blocks per SM each; they exist so that data-parallel buckets can leave early)
        self._unpack_all = L.WeightJobs(all_jobs, self.device)
        self._bwd_ready = True

    # ------------------------------------------------------------------------------------------
    def pack_weights(self):
        """fp32 OIHW / IOHW parameters -> K-major operand packs: one batched launch, only when a parameter changed."""
        params = list(self.model.parameters())
        versions = tuple(p._version for p in params)
        ptrs = tuple(p.data_ptr() for p in params)
        if (versions, ptrs) == self._pack_versions:
            return
        if self._pack_jobs is None or self._pack_ptrs != ptrs:
            jobs = []
            for l in self.layers:
                w = l.conv.weight
                if l.first and self.pair_first:
                    jobs.append((w.data_ptr(), l.wf.data_ptr(), None, 3, l.cout, l.cin, 2 * self.kpad))
                elif l.first:
                    jobs.append((w.data_ptr(), l.wf.data_ptr(), None, 2, l.cout, l.cin, self.kpad))
                else:
                    jobs.append((w.data_ptr(), l.wf.data_ptr(), l.wd.data_ptr(), 0, l.cout, l.cin, 0))
            for ct in self.convts:
                jobs.append((ct.mod.weight.data_ptr(), ct.wf.data_ptr(), ct.wd.data_ptr(), 1, ct.cout, ct.cin, 0))
            self._pack_jobs = L.WeightJobs(jobs, self.device)
            self._pack_ptrs = ptrs
        L.weights_pack(self._pack_jobs, self.dt)
        self._pack_versions = (versions, ptrs)

    # ------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, training: bool) -> torch.Tensor:
        m = self.model
        self.generation += 1
        self.training_pass = training
        self.pack_weights()
        if training:
            self.acc64.zero_()
        L.im2col3x3_first(x, self.xcol)
        n = self.n

        # last block: BatchNorm apply + ReLU fused with the head; its activation is stored only if a later pass reads it
        # (the unfused head backward for dout > 4)
        fuse_head = self.dout <= 4 and os.environ.get("UNETK_FUSE_HEAD", "1") == "1"
        last = self.dec[3][1]
        logits = torch.empty((n, self.dout, self.h, self.w), dtype=torch.float32, device=self.device)

        def conv_bn(l: _ConvBN):
            count = n * l.h * l.w
            L.LABEL = l.name
            if l.first and self.pair_first:
                fs = self.first_stats
                L.conv(self._pairs(l.src), l.wf, self._pairs(l.z), L.MODE_1X1,
                       stat_sum=fs[:128] if training else None, stat_sumsq=fs[128:] if training else None,
                       algo=self.algo, algo_flops=2 * count * 9 * l.cin * l.cout)
                if training:     # the two pixels of a pair are the same 64 BatchNorm channels
                    torch.add(fs[:64], fs[64:128], out=l.stat_sum)
                    torch.add(fs[128:192], fs[192:], out=l.stat_sumsq)
            else:
                L.conv(l.src, l.wf, l.z, L.MODE_1X1 if l.first else L.MODE_3X3,
                       stat_sum=l.stat_sum if training else None, stat_sumsq=l.stat_sumsq if training else None,
                       algo=self.algo, algo_flops=(2 * count * 9 * l.cin * l.cout) if l.first else None)
            bn = l.bn
            if bn.momentum is None:
                # nn.BatchNorm2d(momentum=None) means a cumulative moving average whose factor 1/num_batches_tracked lives
                # on the device; reading it would stall the stream (and break graph capture), so it is refused loudly
                raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative average) is not supported by the fused "
                                          "engine; the reference always uses the default momentum 0.1 (unet/unet.py:17,20)")
            momentum = bn.momentum
            track = bn.track_running_stats and bn.running_mean is not None
            L.bn_finalize(l.stat_sum, l.stat_sumsq, count, l.cout, training, bn.weight, bn.bias, l.conv.bias,
                          bn.running_mean if track else None, bn.running_var if track else None,
                          bn.num_batches_tracked if (track and training) else None,
                          momentum, bn.eps, l.scale, l.shift, l.mean, l.invstd)
            if l is last and fuse_head:
                L.bn_relu_head_fprop(l.z, l.scale, l.shift, None, m.output.weight, m.output.bias, self.dout, logits)
            else:
                L.bn_relu_apply(l.z, l.scale, l.shift, l.a, l.pooled, l.pool_idx)

        for lvl in range(5):
            conv_bn(self.enc[lvl][0])
            conv_bn(self.enc[lvl][1])
        for i in range(4):
            ct = self.convts[i]
            L.LABEL = ct.name
            L.conv(ct.src, ct.wf, ct.out, L.MODE_CONVT, bias=ct.mod.bias, algo=self.algo)
            conv_bn(self.dec[i][0])
            conv_bn(self.dec[i][1])
        if not fuse_head:
            L.head_fprop(self.head_in, m.output.weight, m.output.bias, self.dout, logits)
        return logits

    # ------------------------------------------------------------------------------------------
    def backward(self, dlogits: torch.Tensor, bucket_hook: Optional[Callable] = None) -> Dict[torch.nn.Parameter, torch.Tensor]:
        """Returns {parameter: gradient} with gradients being views of one freshly allocated flat buffer."""
        if not self.training_pass:
            raise RuntimeError("backward through an eval-mode (running statistics) forward is not supported; "
                               "call model.train() before the forward pass")
        if not self._bwd_ready:
            self._build_backward()
        m = self.model
        flat = torch.zeros(self.grad_total, dtype=torch.float32, device=self.device)
        g = {}
        for p, o0, o1 in zip(self.grad_order, self.grad_offsets[:-1], self.grad_offsets[1:]):
            g[p] = flat[o0:o1].view(p.shape)
        self.ws.zero_()
        self.bwd64.zero_()
        self.flat_grad = flat
        dlogits = dlogits.contiguous()

        # head backward: fused with the BatchNorm backward of the last block when the class count allows it
        fuse_head = self.dout <= 4 and os.environ.get("UNETK_FUSE_HEAD", "1") == "1"
        if not fuse_head:
            L.head_bwd(dlogits, self.head_in, m.output.weight, self.dout, self.g_head_in, g[m.output.weight],
                       g[m.output.bias])
        if bucket_hook and not fuse_head:
            bucket_hook(self, self.grad_offsets[2])

        main = torch.cuda.current_stream()
        # per-launch instrumentation (bench.py / tools) needs un-overlapped kernels for meaningful event durations
        side = self.side_stream if L.PROFILE_HOOK is None else None

        def on_side(fn):
            """Run `fn` (weight-gradient launches) on the side stream, ordered after everything enqueued so far."""
            if side is None:
                fn()
                return
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fn()

        def join_side():
            if side is not None:
                main.wait_stream(side)

        fuse = os.environ.get("UNETK_FUSE_BN_REDUCE", "1") == "1"

        def bn_red(l: _ConvBN):
            """Arguments that make a data-gradient launch also accumulate layer l's BatchNorm-backward reductions."""
            return (l.z, l.scale, l.shift, l.mean, l.invstd, l.bwd_sums) if fuse else None

        def conv_bn_bwd(l: _ConvBN, dy, dpool=None, reduced=False, feeds: Optional[_ConvBN] = None, from_head=False):
            """`reduced`: this layer's reductions were already fused into the launch that produced `dy`;
            `feeds`: the layer whose activated-output gradient this layer's dgrad produces (no pooling in between);
            `from_head`: l feeds the classifier head and dy is never materialised (fused head + BatchNorm backward)."""
            L.LABEL = l.name
            if from_head:
                L.head_bn_bwd(dlogits, l.z, m.output.weight, self.dout, l.scale, l.shift, l.mean, l.invstd, self.head_sums,
                              l.dz, g[l.bn.weight], g[l.bn.bias], g[m.output.weight], g[m.output.bias])
                if bucket_hook:
                    bucket_hook(self, self.grad_offsets[2])
            else:
                L.bn_relu_bwd(l.z, dy, dpool, l.scale, l.shift, l.mean, l.invstd, l.bwd_sums, l.dz, g[l.bn.weight],
                              g[l.bn.bias], pool_idx=l.pool_idx if dpool is not None else None, reduced=reduced and fuse)
            if l.first and self.pair_first:
                # dW' [2*cout][2*kpad] over pixel pairs; weights_unpack (kind 3) adds its two diagonal blocks
                on_side(lambda: L.wgrad(self._pairs(l.dz), self._pairs(l.src), l.ws, 0, algo=self.algo,
                                        algo_flops=2 * self.n * l.h * l.w * 9 * l.cin * l.cout))
            elif l.first:
                on_side(lambda: L.wgrad(l.dz, l.src, l.ws, 0, algo=self.algo,
                                        algo_flops=2 * self.n * l.h * l.w * 9 * l.cin * l.cout))
            else:
                # data gradient first on the main stream (the next layer's BatchNorm backward waits for it) ...
                L.conv(l.dz, l.wd, l.g_in, L.MODE_3X3, algo=self.algo, bn_reduce=bn_red(feeds) if feeds is not None else None)
                # ... weight gradient on the side stream
                on_side(lambda: L.wgrad(l.dz, l.src, l.ws, 1, algo=self.algo))
            # conv bias in front of train-mode BN: its gradient is identically zero (flat buffer is zeroed)

        grad_a2 = self.g_head_in
        pos = 2
        seg = 0
        for i in (3, 2, 1, 0):
            lvl = 3 - i
            l1, l2 = self.dec[i]
            # grad_a2 comes from the head (i == 3) or from the ConvTranspose data gradient of the level above, which
            # already accumulated l2's reductions
            conv_bn_bwd(l2, grad_a2, reduced=(i != 3), feeds=l1, from_head=(i == 3 and fuse_head))
            conv_bn_bwd(l1, l2.g_in, reduced=True)
            ct = self.convts[i]
            L.LABEL = ct.name
            nxt = self.dec[i - 1][1] if i > 0 else self.enc[4][1]     # layer whose activated output feeds this ConvTranspose
            L.conv(ct.g_out, ct.wd, ct.g_in, L.MODE_CONVT_GATHER, algo=self.algo, bn_reduce=bn_red(nxt))

            def ct_grads(ct=ct):
                L.wgrad(ct.src, ct.g_out, ct.ws, 2, algo=self.algo)
                L.channel_sum(ct.g_out, g[ct.mod.bias])
            on_side(ct_grads)
            grad_a2 = ct.g_in
            pos += 10
            join_side()
            if bucket_hook:
                L.weights_unpack(self._unpack_jobs[seg], flat)
                bucket_hook(self, self.grad_offsets[pos])
            seg += 1
        for lvl in (4, 3, 2, 1, 0):
            l1, l2 = self.enc[lvl]
            if lvl == 4:
                conv_bn_bwd(l2, grad_a2, reduced=True, feeds=l1)
            else:
                # two gradient sources (skip + max-pool): the stand-alone reduction kernel handles the routing
                conv_bn_bwd(l2, self.dcat[lvl][..., :ENC_CH[lvl]], dpool=self.enc[lvl + 1][0].g_in, feeds=l1)
            conv_bn_bwd(l1, l2.g_in, reduced=True)
            pos += 8
            join_side()
            if bucket_hook:
                L.weights_unpack(self._unpack_jobs[seg], flat)
                bucket_hook(self, self.grad_offsets[pos])
            seg += 1
        if not bucket_hook:
            L.weights_unpack(self._unpack_all, flat)
        return g


class _UNetFunction(torch.autograd.Function):
    """The single autograd node standing for unet.forward (unet/unet.py:93-105)."""

    @staticmethod
    def forward(ctx, model, plan, x, *params):
        ctx.plan = plan
        ctx.model = model
        ctx.params = params
        ctx.generation = plan.generation + 1
        logits = plan.forward(x, training=model.training)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        plan: UNetPlan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("the activations of this forward pass were overwritten by a later forward pass")
        hook = getattr(ctx.model, "_bucket_hook", None)
        scale = getattr(ctx.model, "_grad_scale", None)
        if scale is not None and scale != 1.0:
            dlogits = dlogits * scale
        grads = plan.backward(dlogits, hook)
        plan.busy = False
        done = getattr(ctx.model, "_backward_done_hook", None)
        if done:
            done(plan)
        out = [grads.get(p) if p.requires_grad else None for p in ctx.params]
        return (None, None, None, *out)


class UNetEngine:
    """Plan cache + entry point used by ``unet.forward``."""

    def __init__(self, model):
        self.model = model
        self.plans: Dict[tuple, List[UNetPlan]] = {}

    def run(self, x: torch.Tensor) -> torch.Tensor:
        m = self.model
        params = list(m.parameters())
        L.require_cuda(x, params[0])
        if x.dim() != 4 or x.shape[1] != m.din:
            raise RuntimeError(f"expected input [N,{m.din},H,W], got {tuple(x.shape)}")
        if x.device != params[0].device:
            raise RuntimeError(f"input is on {x.device} but the model is on {params[0].device}")
        for p in params:
            if p.dtype != torch.float32:
                raise RuntimeError("parameters must stay fp32 (master weights); precision is selected with model.precision")
        precision = os.environ.get("UNETK_PRECISION", m.precision)
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        n, _, h, w = x.shape
        key = (n, h, w, precision, x.device.index)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        pool = self.plans.setdefault(key, [])
        plan = next((p for p in pool if not p.busy), None)
        if plan is None and len(pool) >= 2:
            # forwards whose backward never ran (e.g. a validation loss computed with grad enabled) must not
            # leak buffers: recycle the oldest plan; a late backward on it fails the generation check loudly
            plan = min(pool, key=lambda p: p.generation)
        if plan is None:
            with torch.cuda.device(x.device):
                plan = UNetPlan(m, n, h, w, precision, x.device)
            pool.append(plan)
        plan.algo = {"auto": L.ALGO_AUTO, "simt": L.ALGO_SIMT, "tc": L.ALGO_TC}[os.environ.get("UNETK_ALGO", m.conv_algo)]
        with torch.cuda.device(x.device):
            if needs_grad:
                plan.busy = True
                return _UNetFunction.apply(m, plan, x, *params)
            with torch.no_grad():
                return plan.forward(x, training=m.training)
