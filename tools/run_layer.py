"""Run one contraction of the path in isolation (for ncu / quick timing):
    python tools/run_layer.py --op conv --cin 64 --cout 64 --hw 256 --batch 16 [--iters 5]
ops: conv (3x3 fprop with BN statistics), dgrad (3x3 without statistics), wgrad (3x3), convt, convt_dgrad"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from image_segmentation_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="conv")
ap.add_argument("--cin", type=int, default=64)
ap.add_argument("--cout", type=int, default=64)
ap.add_argument("--hw", type=int, default=256)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = "cuda"
n, h, w = a.batch, a.hw, a.hw
bf = torch.bfloat16
x = torch.randn(n, h, w, a.cin, device=dev).to(bf)
if a.op in ("conv", "dgrad", "dgrad_bn"):
    wt = (torch.randn(a.cout, 9, a.cin, device=dev) * 0.05).to(bf)
    y = torch.empty(n, h, w, a.cout, device=dev, dtype=bf)
    s1 = torch.zeros(a.cout, dtype=torch.float64, device=dev)
    s2 = torch.zeros_like(s1)
    flops = 2 * n * h * w * 9 * a.cin * a.cout

    zt = torch.randn(n, h, w, a.cout, device=dev).to(bf)
    vec = [torch.rand(a.cout, device=dev) + 0.5 for _ in range(4)]
    sums = torch.zeros(2 * a.cout, dtype=torch.float64, device=dev)

    def run():
        if a.op == "conv":
            L.conv(x, wt, y, L.MODE_3X3, stat_sum=s1, stat_sumsq=s2)
        elif a.op == "dgrad_bn":
            L.conv(x, wt, y, L.MODE_3X3, bn_reduce=(zt, vec[0], vec[1], vec[2], vec[3], sums))
        else:
            L.conv(x, wt, y, L.MODE_3X3)
elif a.op == "wgrad":
    dy = torch.randn(n, h, w, a.cout, device=dev).to(bf)
    dw = torch.zeros(a.cout, 9, a.cin, device=dev)
    flops = 2 * n * h * w * 9 * a.cin * a.cout

    def run():
        L.wgrad(dy, x, dw, 1)
elif a.op == "convt":
    wt = (torch.randn(4 * a.cout, a.cin, device=dev) * 0.05).to(bf)
    y = torch.empty(n, 2 * h, 2 * w, a.cout, device=dev, dtype=bf)
    b = torch.zeros(a.cout, device=dev)
    flops = 2 * n * h * w * a.cin * 4 * a.cout

    def run():
        L.conv(x, wt, y, L.MODE_CONVT, bias=b)
elif a.op == "convt_dgrad":
    # data gradient of ConvTranspose2d(cin -> cout): gathers the hi-res gradient [N,2h,2w,cout] into [N,h,w,cin]
    dyh = torch.randn(n, 2 * h, 2 * w, a.cout, device=dev).to(bf)
    wt = (torch.randn(a.cin, 4, a.cout, device=dev) * 0.05).to(bf)
    y = torch.empty(n, h, w, a.cin, device=dev, dtype=bf)
    flops = 2 * n * h * w * a.cin * 4 * a.cout

    def run():
        L.conv(dyh, wt, y, L.MODE_CONVT_GATHER)
elif a.op == "bilinear_fwd" or a.op == "bilinear_bwd":
    # clip/clipunet.py:99-100 at the last decoder level by default: 14x14 -> hw x hw, cin channels
    src = torch.randn(n, 14, 14, a.cin, device=dev).to(bf)
    dst = torch.randn(n, h, w, a.cin, device=dev).to(bf)
    flops = 0

    def run():
        L.bilinear_up(src, dst, backward=a.op.endswith("bwd"))
else:
    raise SystemExit("unknown op")
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"{a.op} cin={a.cin} cout={a.cout} hw={a.hw} batch={a.batch}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
if a.op.startswith("bilinear"):
    print(f"  bytes (destination tensor once): {dst.numel() * 2 / 1e6:.1f} MB -> {dst.numel() * 2 / ms / 1e6:.0f} GB/s")
if os.environ.get("UNETK_DBG") == "1":
    import ctypes
    buf = (ctypes.c_longlong * (148 * 8))()
    L.lib().unetk_debug_counters(buf, 148 * 8)     # only in builds with UNETK_NVCC_EXTRA=-DUNETK_DEBUG_COUNTERS
    import numpy as np
    d = np.array(buf[:]).reshape(148, 8)
    if not (L.TC_FLAGS & (L.TC_FLAG_BITS["no_halo_pair"] | L.TC_FLAG_BITS["no_pair"])):
        d = d[::2]     # leader CTAs hold the MMA counters
    names = ["prod wait emptyA", "mma wait fullA", "mma wait fullB", "mma wait tmem_empty", "mma total",
             "epi0 wait tmem_full", "epi0 in epilogue", "epi0 total"]
    for i, nm in enumerate(names):
        print(f"{nm:22s} mean {d[:, i].mean():12.0f}  min {d[:, i].min():12d}  max {d[:, i].max():12d} cycles")
