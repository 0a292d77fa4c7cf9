"""Time one bandwidth kernel of the path in isolation: python tools/run_elem.py [--op head_fwd|bn_apply|head_bwd] [--dout 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from image_segmentation_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="head_fwd")
ap.add_argument("--dout", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--hw", type=int, default=256)
ap.add_argument("--c", type=int, default=64)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev, bf = "cuda", torch.bfloat16
n, h, w, c = a.batch, a.hw, a.hw, a.c
z = torch.randn(n, h, w, c, device=dev).to(bf)
z2 = torch.randn(n, h, w, c, device=dev).to(bf)      # second buffer so successive launches do not hit L2
scale, shift = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
mean, invstd = torch.randn(c, device=dev) * 0.1, torch.rand(c, device=dev) + 0.5
wh, bh = torch.randn(a.dout, c, device=dev) * 0.1, torch.randn(a.dout, device=dev)
logits = torch.empty(n, a.dout, h, w, device=dev)
dl = torch.randn(n, a.dout, h, w, device=dev)
act = torch.empty_like(z)
dz = torch.empty_like(z)
sums = torch.zeros((3 + a.dout) * c, dtype=torch.float64, device=dev)
gam, bet = torch.empty(c, device=dev), torch.empty(c, device=dev)
dwh, dbh = torch.empty(a.dout, c, device=dev), torch.empty(a.dout, device=dev)


def run(src):
    if a.op == "head_fwd":
        L.bn_relu_head_fprop(src, scale, shift, None, wh, bh, a.dout, logits)
        return src.numel() * 2 + logits.numel() * 4
    if a.op == "bn_apply":
        L.bn_relu_apply(src, scale, shift, act, None, None)
        return src.numel() * 4
    if a.op == "head_bwd":
        L.head_bn_bwd(dl, src, wh, a.dout, scale, shift, mean, invstd, sums, dz, gam, bet, dwh, dbh)
        return src.numel() * 6 + 2 * dl.numel() * 4
    raise SystemExit("unknown op")


for _ in range(3):
    run(z), run(z2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.iters):
    nbytes = run(z if i % 2 == 0 else z2)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"{a.op} dout={a.dout} [{n},{h},{w},{c}]: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s")
