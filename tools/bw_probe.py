"""Bandwidth probes for the BN kernels: contiguous vs channel-slice (concat buffer) operands."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from image_segmentation_b200 import _lib as L  # noqa: E402

dev = "cuda"
n, h, w, c = 64, 256, 256, 64
bf = torch.bfloat16
z = torch.randn(n, h, w, c, device=dev).to(bf)
scale = torch.rand(c, device=dev) + 0.5
shift = torch.randn(c, device=dev) * 0.1
mean = torch.zeros(c, device=dev)
invstd = torch.ones(c, device=dev)
cat = torch.empty(n, h, w, 2 * c, device=dev, dtype=bf)
dcat = torch.randn(n, h, w, 2 * c, device=dev).to(bf)
a_c = torch.empty(n, h, w, c, device=dev, dtype=bf)
dy_c = torch.randn(n, h, w, c, device=dev).to(bf)
dz = torch.empty_like(z)
pooled = torch.empty(n, h // 2, w // 2, c, device=dev, dtype=bf)
pidx = torch.zeros(n, h // 2, w // 2, c // 8, device=dev, dtype=torch.int16)
dp = torch.randn(n, h // 2, w // 2, c, device=dev).to(bf)
sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)


def timeit(name, fn, gbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:55s} {ms:7.3f} ms  {gbytes / ms:8.1f} GB/s")


E = n * h * w * c * 2 / 1e6  # MB of one 64-channel bf16 tensor
timeit("bn_apply   z -> a (contiguous)", lambda: L.bn_relu_apply(z, scale, shift, a_c), 2 * E)
timeit("bn_apply   z -> a (slice of concat buffer)", lambda: L.bn_relu_apply(z, scale, shift, cat[..., :c]), 2 * E)
timeit("bn_apply+pool z -> a (contiguous), pooled, idx", lambda: L.bn_relu_apply(z, scale, shift, a_c, pooled, pidx), 2.25 * E)
timeit("bn_apply+pool z -> a (slice), pooled, idx", lambda: L.bn_relu_apply(z, scale, shift, cat[..., :c], pooled, pidx), 2.25 * E)
red = lambda dy, dpool: (L.lib().unetk_bn_relu_bwd_reduce, None)


def bwd(dy, dpool):
    L.bn_relu_bwd(z, dy, dpool, scale, shift, mean, invstd, sums, dz, dg, db, pool_idx=pidx if dpool is not None else None)


timeit("bn_bwd (reduce+apply) dy contiguous", lambda: bwd(dy_c, None), 5 * E)
timeit("bn_bwd (reduce+apply) dy slice", lambda: bwd(dcat[..., :c], None), 5 * E)
timeit("bn_bwd (reduce+apply) dy contiguous + dpool", lambda: bwd(dy_c, dp), 5.5 * E)
timeit("bn_bwd (reduce+apply) dy slice + dpool", lambda: bwd(dcat[..., :c], dp), 5.5 * E)
x = torch.empty(n * h * w * c, device=dev, dtype=bf)
y = torch.empty_like(x)
timeit("torch copy_ (reference)", lambda: y.copy_(x), 2 * E)
