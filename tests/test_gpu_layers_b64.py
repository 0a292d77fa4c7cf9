"""Per-layer parity of the tcgen05 contraction kernels AT THE BENCHMARKED CONFIGURATION (batch 64, 256x256):
every one of the 23 contraction layers of unet(3,3) (SURVEY.md section 8(a) layer table; reference
unet/unet.py:16,19,59,91) x {forward (+ BatchNorm batch statistics), data gradient (+ fused BatchNorm-backward sums),
weight gradient}, called through the C ABI with UNETK_ALGO_TC exactly as the engine calls them, against an independent
float64 reference.

Why this file exists: several code paths only trigger at this size -- the grouped PairSched schedule and the trimmed pair
count (Cout >= 512 with >= 74 pair tiles), resident weights with > 200 tiles per persistent CTA (mbarrier phase wrap),
split-K weight gradients over 4.2 M pixels, the nine-tap 64-channel weight-gradient kernel at 256^2.

Reference: torch.nn.functional conv2d / conv_transpose2d in float64 on the GPU (cuDNN / ATen -- none of this repo's
kernels), itself pinned against CPU float64 on a slice of the batch (`_pin_reference`).  Inputs are exactly
representable in bf16, so products are exact and the only error sources are fp32 accumulation order and the bf16
rounding of the stored output.

Tolerances (north_star: bf16 tier rel 2e-2): activations element-wise |err| <= 1 bf16 ulp (2^-8 |ref|) + 1e-5 max|ref|
-- far inside the 2e-2 budget; weight gradients max-abs 1e-3 relative; fused per-channel sums 1e-5 relative to the
sum of magnitudes (they are accumulated from the STORED bf16 values, so they are compared with float64 sums of the
kernel's own output).
"""
import zlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from image_segmentation_b200 import _lib as L  # noqa: E402

DEV = "cuda"
N = 64
BF = torch.bfloat16

# (name, cin, cout, hw) of the 17 ordinary 3x3 layers (the first conv, Cin = 3, is tested separately)
CONV3 = [("down1.c2", 64, 64, 256), ("down2.c1", 64, 128, 128), ("down2.c2", 128, 128, 128),
         ("down3.c1", 128, 256, 64), ("down3.c2", 256, 256, 64), ("down4.c1", 256, 512, 32), ("down4.c2", 512, 512, 32),
         ("down5.c1", 512, 1024, 16), ("down5.c2", 1024, 1024, 16), ("up1.c1", 1024, 512, 32), ("up1.c2", 512, 512, 32),
         ("up2.c1", 512, 256, 64), ("up2.c2", 256, 256, 64), ("up3.c1", 256, 128, 128), ("up3.c2", 128, 128, 128),
         ("up4.c1", 128, 64, 256), ("up4.c2", 64, 64, 256)]
CONVT = [("up1.upsample", 1024, 512, 16), ("up2.upsample", 512, 256, 32), ("up3.upsample", 256, 128, 64),
         ("up4.upsample", 128, 64, 128)]


def rnd(shape, seed, scale=1.0, dtype=BF):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, generator=g, device=DEV) * scale).to(dtype)


def nchw64(t):
    """NHWC (any dtype) -> NCHW float64 view-copy."""
    return t.permute(0, 3, 1, 2).double()


def assert_activation(y, ref_nhwc, what):
    """|err| <= one bf16 ulp of the reference + a sliver of the tensor's max (fp32 accumulation noise near zero)."""
    ref = ref_nhwc
    err = (y.double() - ref).abs()
    bound = ref.abs() * 2.0 ** -8 + 1e-5 * ref.abs().max()
    bad = int((err > bound).sum())
    rel = (err.max() / ref.abs().max()).item()
    assert bad == 0, f"{what}: {bad} elements beyond 1 bf16 ulp (max-abs rel {rel:.3e})"
    assert rel < 2e-2        # the north-star budget, for the record
    return rel


def assert_sums(got, ref, mag, what, tol=1e-5):
    err = (got.double() - ref).abs()
    assert bool((err <= tol * mag + 1e-12).all()), f"{what}: max err {err.max().item():.3e} vs magnitude {mag.max().item():.3e}"


def _pin_reference(fn_gpu, fn_cpu, what):
    """The GPU float64 reference must agree with CPU float64 (1e-10) on the slice both compute."""
    a, b = fn_gpu().cpu(), fn_cpu()
    assert (a - b).abs().max().item() <= 1e-10 * max(1.0, b.abs().max().item()), what


def bn_vectors(c, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    scale = torch.rand(c, generator=g, device=DEV) + 0.5
    shift = torch.randn(c, generator=g, device=DEV) * 0.5
    mean = torch.randn(c, generator=g, device=DEV) * 0.3
    invstd = torch.rand(c, generator=g, device=DEV) + 0.5
    return scale, shift, mean, invstd


def bn_bwd_sums_ref(dA, z, scale, shift, mean, invstd):
    """float64 sums over the STORED dA: dy = dA * [relu(z*scale+shift) > 0]; s1 = sum dy; s2 = sum dy * xhat."""
    # the kernel evaluates fmaf(z, scale, shift) > 0 in fp32: the sign of a fused multiply-add is the sign of the exact
    # value, which float64 reproduces (bf16 x fp32 products are exact in float64)
    mask = (z.double() * scale.double() + shift.double()) > 0
    dy = dA.double() * mask
    xhat = (z.double() - mean.double()) * invstd.double()
    s1 = dy.sum(dim=(0, 1, 2))
    s2 = (dy * xhat).sum(dim=(0, 1, 2))
    mag1 = dy.abs().sum(dim=(0, 1, 2))
    mag2 = (dy * xhat).abs().sum(dim=(0, 1, 2)) + mean.double().abs() * invstd.double() * mag1
    return s1, s2, mag1, mag2


@pytest.fixture(autouse=True)
def _free_memory():
    yield
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name,cin,cout,hw", CONV3, ids=[c[0] for c in CONV3])
def test_conv3x3_layer_at_batch_64(name, cin, cout, hw):
    seed = zlib.crc32(name.encode()) % 10000
    x = rnd((N, hw, hw, cin), seed)
    wt = rnd((cout, cin, 3, 3), seed + 1, 0.05)                       # parameter layout, bf16-exact values
    wf = wt.permute(0, 2, 3, 1).contiguous()                          # [co][t][ci]
    wd = wt.flip(2, 3).permute(1, 2, 3, 0).contiguous()               # [ci][8-t][co]
    x64, w64 = nchw64(x), wt.double()

    # ---- forward + BatchNorm batch statistics (unet/unet.py:16-17) ----
    y = torch.empty((N, hw, hw, cout), dtype=BF, device=DEV)
    s1 = torch.zeros(cout, dtype=torch.float64, device=DEV)
    s2 = torch.zeros_like(s1)
    L.conv(x, wf, y, L.MODE_3X3, stat_sum=s1, stat_sumsq=s2, algo=L.ALGO_TC)
    ref = F.conv2d(x64, w64, padding=1).permute(0, 2, 3, 1)
    _pin_reference(lambda: F.conv2d(x64[-1:], w64, padding=1), lambda: F.conv2d(x64[-1:].cpu(), w64.cpu(), padding=1), name + " fprop ref")
    assert_activation(y, ref, name + " fprop")
    yd = y.double()
    assert_sums(s1, yd.sum(dim=(0, 1, 2)), yd.abs().sum(dim=(0, 1, 2)), name + " sum y")
    assert_sums(s2, (yd * yd).sum(dim=(0, 1, 2)), (yd * yd).sum(dim=(0, 1, 2)), name + " sum y^2")
    del ref, yd

    # ---- data gradient with the fused BatchNorm-backward reduction of the producer layer (autograd of :16,19) ----
    dy = rnd((N, hw, hw, cout), seed + 2)
    z = rnd((N, hw, hw, cin), seed + 3, 2.0)
    scale, shift, mean, invstd = bn_vectors(cin, seed + 4)
    sums = torch.zeros(2 * cin, dtype=torch.float64, device=DEV)
    dx = torch.empty((N, hw, hw, cin), dtype=BF, device=DEV)
    L.conv(dy, wd, dx, L.MODE_3X3, algo=L.ALGO_TC, bn_reduce=(z, scale, shift, mean, invstd, sums))
    dy64 = nchw64(dy)
    ref_dx = F.conv_transpose2d(dy64, w64, padding=1).permute(0, 2, 3, 1)
    _pin_reference(lambda: F.conv_transpose2d(dy64[:1], w64, padding=1),
                   lambda: F.conv_transpose2d(dy64[:1].cpu(), w64.cpu(), padding=1), name + " dgrad ref")
    assert_activation(dx, ref_dx, name + " dgrad")
    r1, r2, m1, m2 = bn_bwd_sums_ref(dx, z, scale, shift, mean, invstd)
    assert_sums(sums[:cin], r1, m1, name + " bn-bwd sum dy", tol=2e-5)
    assert_sums(sums[cin:], r2, m2, name + " bn-bwd sum dy*xhat", tol=2e-5)
    del ref_dx, z, dx

    # ---- weight gradient: split-K over N*H*W pixels, fp32 atomics ----
    dw = torch.zeros((cout, 9, cin), dtype=torch.float32, device=DEV)
    L.wgrad(dy, x, dw, 1, algo=L.ALGO_TC)
    ref_dw = torch.nn.grad.conv2d_weight(x64, (cout, cin, 3, 3), dy64, padding=1)
    _pin_reference(lambda: torch.nn.grad.conv2d_weight(x64[:1], (cout, cin, 3, 3), dy64[:1], padding=1),
                   lambda: torch.nn.grad.conv2d_weight(x64[:1].cpu(), (cout, cin, 3, 3), dy64[:1].cpu(), padding=1),
                   name + " wgrad ref")
    ref_dw = ref_dw.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    rel = ((dw.double() - ref_dw).abs().max() / ref_dw.abs().max()).item()
    assert rel < 1e-3, f"{name} wgrad: max-abs rel {rel:.3e}"


@pytest.mark.parametrize("name,cin,cout,hw", CONVT, ids=[c[0] for c in CONVT])
def test_conv_transpose_layer_at_batch_64(name, cin, cout, hw):
    seed = zlib.crc32(name.encode()) % 10000
    x = rnd((N, hw, hw, cin), seed)
    wt = rnd((cin, cout, 2, 2), seed + 1, 0.05)
    b = rnd((cout,), seed + 2, 1.0, torch.float32)
    wf = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous()   # [(a,b,co)][ci]
    wd = wt.permute(0, 2, 3, 1).reshape(cin, 4, cout).contiguous()    # [ci][(a,b)][co]
    x64, w64 = nchw64(x), wt.double()
    # forward into the second half of the concat buffer (unet/unet.py:59,63)
    cat = torch.full((N, 2 * hw, 2 * hw, 2 * cout), 7.0, dtype=BF, device=DEV)
    out = cat[..., cout:]
    L.conv(x, wf, out, L.MODE_CONVT, bias=b, algo=L.ALGO_TC)
    ref = F.conv_transpose2d(x64, w64, b.double(), stride=2).permute(0, 2, 3, 1)
    _pin_reference(lambda: F.conv_transpose2d(x64[:1], w64, b.double(), stride=2),
                   lambda: F.conv_transpose2d(x64[:1].cpu(), w64.cpu(), b.double().cpu(), stride=2), name + " fprop ref")
    assert_activation(out, ref, name + " fprop")
    assert bool((cat[..., :cout] == 7.0).all()), "skip half of the concat buffer was overwritten"
    del ref, cat, out
    # data gradient (4 stride-2 gather maps) with the fused BatchNorm-backward reduction
    dcat = torch.zeros((N, 2 * hw, 2 * hw, 2 * cout), dtype=BF, device=DEV)
    dy = dcat[..., cout:]
    dy.copy_(rnd((N, 2 * hw, 2 * hw, cout), seed + 3))
    z = rnd((N, hw, hw, cin), seed + 4, 2.0)
    scale, shift, mean, invstd = bn_vectors(cin, seed + 5)
    sums = torch.zeros(2 * cin, dtype=torch.float64, device=DEV)
    dx = torch.empty((N, hw, hw, cin), dtype=BF, device=DEV)
    L.conv(dy, wd, dx, L.MODE_CONVT_GATHER, algo=L.ALGO_TC, bn_reduce=(z, scale, shift, mean, invstd, sums))
    dy64 = nchw64(dy)
    ref_dx = F.conv2d(dy64, w64, stride=2).permute(0, 2, 3, 1)
    assert_activation(dx, ref_dx, name + " dgrad")
    r1, r2, m1, m2 = bn_bwd_sums_ref(dx, z, scale, shift, mean, invstd)
    assert_sums(sums[:cin], r1, m1, name + " bn-bwd sum dy", tol=2e-5)
    assert_sums(sums[cin:], r2, m2, name + " bn-bwd sum dy*xhat", tol=2e-5)
    del ref_dx, dx, z
    # weight gradient (mode 2) and bias gradient
    dw = torch.zeros((cin, 4, cout), dtype=torch.float32, device=DEV)
    L.wgrad(x, dy, dw, 2, algo=L.ALGO_TC)
    ref_dw = torch.einsum("nchw,ndhawb->cabd", x64, dy64.reshape(N, cout, hw, 2, hw, 2)).reshape(cin, 4, cout)
    rel = ((dw.double() - ref_dw).abs().max() / ref_dw.abs().max()).item()
    assert rel < 1e-3, f"{name} wgrad: max-abs rel {rel:.3e}"
    db = torch.zeros(cout, dtype=torch.float32, device=DEV)
    L.channel_sum(dy, db)
    ref_db = dy64.sum(dim=(0, 2, 3))
    assert ((db.double() - ref_db).abs().max() / dy64.abs().sum(dim=(0, 2, 3)).max()).item() < 1e-5


def test_first_layer_pixel_pair_form_at_batch_64():
    """down1.doubleConvReLU.0 (Cin = 3) as the engine runs it: im2col (K = 32 per pixel) -> pixel-pair 1x1 GEMM with
    the block-diagonal weight -> statistics arriving as two halves; weight gradient = sum of the diagonal blocks."""
    H = W = 256
    g = torch.Generator(device=DEV).manual_seed(7)
    x = torch.rand((N, 3, H, W), generator=g, device=DEV)
    wt = (torch.randn((64, 3, 3, 3), generator=g, device=DEV) * 0.2)
    kpad = 32
    xcol = torch.empty((N, H, W, kpad), dtype=BF, device=DEV)
    L.im2col3x3_first(x, xcol)
    wf = torch.zeros((128, 2 * kpad), dtype=BF, device=DEV)
    L.weights_pack(L.WeightJobs([(wt.data_ptr(), wf.data_ptr(), None, 3, 64, 3, 2 * kpad)], DEV), BF)
    z = torch.empty((N, H, W, 64), dtype=BF, device=DEV)
    fs = torch.zeros(256, dtype=torch.float64, device=DEV)
    pairs = lambda t: t.view(t.shape[0], t.shape[1], t.shape[2] // 2, 2 * t.shape[3])  # noqa: E731
    L.conv(pairs(xcol), wf, pairs(z), L.MODE_1X1, stat_sum=fs[:128], stat_sumsq=fs[128:], algo=L.ALGO_TC)
    xb, wb = x.to(BF).double(), wt.to(BF).double()                     # what the kernels see: bf16-rounded operands
    ref = F.conv2d(xb, wb, padding=1).permute(0, 2, 3, 1)
    _pin_reference(lambda: F.conv2d(xb[:2], wb, padding=1), lambda: F.conv2d(xb[:2].cpu(), wb.cpu(), padding=1), "first conv ref")
    assert_activation(z, ref, "down1.c1 fprop")
    zd = z.double()
    assert_sums(fs[:64] + fs[64:128], zd.sum(dim=(0, 1, 2)), zd.abs().sum(dim=(0, 1, 2)), "down1.c1 sum")
    assert_sums(fs[128:192] + fs[192:], (zd * zd).sum(dim=(0, 1, 2)), (zd * zd).sum(dim=(0, 1, 2)), "down1.c1 sumsq")
    # weight gradient over the pair operand; unpack (kind 3) folds the two diagonal blocks
    dz = rnd((N, H, W, 64), 11)
    ws = torch.zeros(128 * 2 * kpad, dtype=torch.float32, device=DEV)
    L.wgrad(pairs(dz), pairs(xcol), ws, 0, algo=L.ALGO_TC)
    flat = torch.zeros(64 * 27, dtype=torch.float32, device=DEV)
    L.weights_unpack(L.WeightJobs([(ws.data_ptr(), 0, None, 3, 64, 3, 2 * kpad)], DEV), flat)
    ref_dw = torch.nn.grad.conv2d_weight(xb, (64, 3, 3, 3), nchw64(dz), padding=1)
    rel = ((flat.view(64, 3, 3, 3).double() - ref_dw).abs().max() / ref_dw.abs().max()).item()
    assert rel < 1e-3, rel


@pytest.mark.parametrize("dout", [3, 4, 1])
def test_fused_head_passes_at_batch_64(dout):
    """output 1x1 head (unet/unet.py:91) fused with the last block's BatchNorm+ReLU (forward) and BatchNorm backward."""
    H = W = 256
    C = 64
    z = rnd((N, H, W, C), 21, 2.0)
    scale, shift, mean, invstd = bn_vectors(C, 22)
    wh = rnd((dout, C), 23, 0.2, torch.float32)
    bh = rnd((dout,), 24, 1.0, torch.float32)
    logits = torch.empty((N, dout, H, W), dtype=torch.float32, device=DEV)
    L.bn_relu_head_fprop(z, scale, shift, None, wh, bh, dout, logits)
    pre = z.double() * scale.double() + shift.double()                 # exact value of the kernel's fp32 fma
    # the fused kernels reproduce the unfused pipeline, in which the activation is STORED as bf16 before the head reads it
    a = pre.clamp_min(0).float().to(BF).double()
    ref = torch.einsum("nhwc,kc->nkhw", a, wh.double()) + bh.double()[None, :, None, None]
    rel = ((logits.double() - ref).abs().max() / ref.abs().max()).item()
    assert rel < 1e-5, f"fused head forward rel {rel:.3e}"               # fp32 arithmetic: north-star fp32 budget 1e-4
    del ref
    # backward: da = dlogits . Wh; dy = da * mask; sums; dz = scale*(dy - s1/M - xhat*s2/M); dWh, dbh, dgamma, dbeta
    dl = rnd((N, dout, H, W), 25, 1.0, torch.float32)
    sums = torch.zeros((3 + dout) * C, dtype=torch.float64, device=DEV)
    dz = torch.empty((N, H, W, C), dtype=BF, device=DEV)
    dgamma, dbeta = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dwh, dbh = torch.zeros((dout, C), device=DEV), torch.zeros(dout, device=DEV)
    L.head_bn_bwd(dl, z, wh, dout, scale, shift, mean, invstd, sums, dz, dgamma, dbeta, dwh, dbh)
    mask = (pre > 0)
    da = torch.einsum("nkhw,kc->nhwc", dl.double(), wh.double())
    dyv = da * mask
    xhat = (z.double() - mean.double()) * invstd.double()
    M = N * H * W
    s1, s2 = dyv.sum(dim=(0, 1, 2)), (dyv * xhat).sum(dim=(0, 1, 2))
    ref_dz = scale.double() * (dyv - s1 / M - xhat * (s2 / M))
    err = (dz.double() - ref_dz).abs()
    assert int((err > ref_dz.abs() * 2.0 ** -8 + 2e-5 * ref_dz.abs().max()).sum()) == 0
    # (the backward pass recomputes a = relu(z*scale+shift) without the bf16 rounding of the stored activation)
    ref_dwh = torch.einsum("nkhw,nhwc->kc", dl.double(), pre.clamp_min(0))
    assert ((dwh.double() - ref_dwh).abs().max() / ref_dwh.abs().max()).item() < 1e-5
    assert ((dbh.double() - dl.double().sum(dim=(0, 2, 3))).abs().max() / dl.double().abs().sum(dim=(0, 2, 3)).max()).item() < 1e-5
    assert ((dgamma.double() - s2).abs().max() / (dyv * xhat).abs().sum(dim=(0, 1, 2)).max()).item() < 1e-5
    assert ((dbeta.double() - s1).abs().max() / dyv.abs().sum(dim=(0, 1, 2)).max()).item() < 1e-5
