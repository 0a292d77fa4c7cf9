"""Per-kernel parity tests of libunetk.so (through the C ABI) against CPU references:
torch.nn.functional in float64 for the floating-point ops (same operator the reference calls),
oracle/ for loss and metrics, golden vectors from the unmodified reference where they exist.

Tolerances (stated per north_star): fp32 tier rel 1e-4, bf16 tier rel 2e-2 (here: max-abs error
relative to the tensor's max-abs); integer outputs bit-exact.
"""
import json

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from image_segmentation_b200 import _lib as L  # noqa: E402
from oracle import loss_oracle, metrics_oracle  # noqa: E402

DEV = "cuda"
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def rnd(shape, dt, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = (torch.randn(shape, generator=g) * scale).to(dt)   # values exactly representable in dt
    return t


def nhwc_to_nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


ALGOS = [(torch.float32, L.ALGO_SIMT), (torch.bfloat16, L.ALGO_SIMT), (torch.bfloat16, L.ALGO_TC)]
IDS = ["f32-simt", "bf16-simt", "bf16-tc"]


@pytest.mark.parametrize("dt,algo", ALGOS, ids=IDS)
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 64), (1, 8, 24, 128, 64), (3, 4, 4, 64, 256), (2, 32, 32, 64, 128),
                                              (5, 2, 2, 128, 128),
                                              # halo kernel: resident 18-slot B, streamed B (K steps > 24), partial tiles,
                                              # several N tiles, many tiles per CTA
                                              (1, 24, 20, 128, 64), (2, 16, 8, 256, 128), (1, 32, 16, 64, 192), (3, 48, 40, 64, 64),
                                              (2, 128, 128, 64, 64), (1, 16, 16, 512, 256)])
def test_conv3x3_fprop_with_stats(dt, algo, n, h, w, cin, cout):
    x = rnd((n, h, w, cin), dt, 1)
    wt = rnd((cout, cin, 3, 3), dt, 2, 0.05)
    ref = F.conv2d(nhwc_to_nchw(x.double()), wt.double(), padding=1).permute(0, 2, 3, 1)
    xd, y = x.to(DEV), torch.empty((n, h, w, cout), dtype=dt, device=DEV)
    wp = wt.permute(0, 2, 3, 1).contiguous().to(DEV)          # [co][r][s][ci]
    s1 = torch.zeros(cout, dtype=torch.float64, device=DEV)
    s2 = torch.zeros(cout, dtype=torch.float64, device=DEV)
    L.conv(xd, wp, y, L.MODE_3X3, stat_sum=s1, stat_sumsq=s2, algo=algo)
    torch.cuda.synchronize()
    assert relerr(y, ref) < TOL[dt]
    yd = y.double().cpu()
    np.testing.assert_allclose(s1.cpu().numpy(), yd.sum(dim=(0, 1, 2)).numpy(), rtol=1e-5, atol=1e-4 * max(1.0, yd.abs().max().item()))
    np.testing.assert_allclose(s2.cpu().numpy(), (yd * yd).sum(dim=(0, 1, 2)).numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dt,algo", ALGOS, ids=IDS)
def test_conv1x1_and_bias(dt, algo):
    n, h, w, cin, cout = 2, 8, 16, 64, 64
    x = rnd((n, h, w, cin), dt, 3)
    wt = rnd((cout, cin), dt, 4, 0.1)
    b = rnd((cout,), torch.float32, 5)
    ref = x.double() @ wt.double().t() + b.double()
    y = torch.empty((n, h, w, cout), dtype=dt, device=DEV)
    L.conv(x.to(DEV), wt.to(DEV), y, L.MODE_1X1, bias=b.to(DEV), algo=algo)
    assert relerr(y, ref) < TOL[dt]


@pytest.mark.parametrize("dt,algo", ALGOS, ids=IDS)
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 4, 4, 128, 64), (1, 8, 8, 512, 256), (3, 2, 2, 64, 64)])
def test_conv_transpose_fprop_into_concat_slice_and_dgrad(dt, algo, n, h, w, cin, cout):
    x = rnd((n, h, w, cin), dt, 6)
    wt = rnd((cin, cout, 2, 2), dt, 7, 0.1)
    b = rnd((cout,), torch.float32, 8)
    ref = F.conv_transpose2d(nhwc_to_nchw(x.double()), wt.double(), b.double(), stride=2).permute(0, 2, 3, 1)
    cat = torch.full((n, 2 * h, 2 * w, 2 * cout), 7.0, dtype=dt, device=DEV)
    out = cat[..., cout:]                                        # second half of the concat buffer
    wf = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous().to(DEV)   # [(a,b,co)][ci]
    L.conv(x.to(DEV), wf, out, L.MODE_CONVT, bias=b.to(DEV), algo=algo)
    assert relerr(out, ref) < TOL[dt]
    assert torch.all(cat[..., :cout] == 7.0)                     # skip half untouched
    # data gradient: dx[n,i,j,ci] = sum dy[n,2i+a,2j+b,co] w[ci,co,a,b]
    dy = rnd((n, 2 * h, 2 * w, cout), dt, 9)
    ref_dx = F.conv2d(nhwc_to_nchw(dy.double()), wt.double().permute(0, 1, 2, 3), stride=2).permute(0, 2, 3, 1)
    dcat = torch.zeros((n, 2 * h, 2 * w, 2 * cout), dtype=dt, device=DEV)
    dcat[..., cout:] = dy.to(DEV)
    wd = wt.permute(0, 2, 3, 1).reshape(cin, 4, cout).contiguous().to(DEV)    # [ci][(a,b)][co]
    dx = torch.empty((n, h, w, cin), dtype=dt, device=DEV)
    L.conv(dcat[..., cout:], wd, dx, L.MODE_CONVT_GATHER, algo=algo)
    assert relerr(dx, ref_dx) < TOL[dt]


@pytest.mark.parametrize("dt,algo", ALGOS, ids=IDS)
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 64), (2, 8, 8, 128, 64), (1, 16, 32, 64, 128), (3, 4, 4, 256, 128),
                                              (2, 32, 24, 128, 256), (1, 48, 40, 64, 64), (4, 64, 64, 64, 64), (1, 16, 16, 512, 128)])
def test_conv3x3_dgrad_and_wgrad(dt, algo, n, h, w, cin, cout):
    x = rnd((n, h, w, cin), dt, 10)
    wt = rnd((cout, cin, 3, 3), dt, 11, 0.05)
    dy = rnd((n, h, w, cout), dt, 12)
    xr = nhwc_to_nchw(x.double()).requires_grad_(True)
    wr = wt.double().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(nhwc_to_nchw(dy.double()))
    # dgrad = 3x3 contraction with the flipped / transposed pack [ci][8-t][co]
    wd = wt.flip(2, 3).permute(1, 2, 3, 0).contiguous().to(DEV)
    dx = torch.empty((n, h, w, cin), dtype=dt, device=DEV)
    L.conv(dy.to(DEV), wd, dx, L.MODE_3X3, algo=algo)
    assert relerr(dx, xr.grad.permute(0, 2, 3, 1)) < TOL[dt]
    dw = torch.zeros((cout, 9, cin), dtype=torch.float32, device=DEV)
    L.wgrad(dy.to(DEV), x.to(DEV), dw, 1, algo=algo)
    ref_dw = wr.grad.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    assert relerr(dw, ref_dw) < 1e-4       # fp32 accumulation of exactly representable inputs
    # accumulation semantics: a second call adds
    L.wgrad(dy.to(DEV), x.to(DEV), dw, 1, algo=algo)
    assert relerr(dw, 2 * ref_dw) < 1e-4


@pytest.mark.parametrize("dt,algo", ALGOS, ids=IDS)
def test_wgrad_1x1_and_convT(dt, algo):
    n, h, w, cin, cout = 2, 8, 8, 128, 64
    x = rnd((n, h, w, cin), dt, 13)
    dy = rnd((n, 2 * h, 2 * w, cout), dt, 14)
    wr = torch.zeros(cin, cout, 2, 2, dtype=torch.float64, requires_grad=True)
    F.conv_transpose2d(nhwc_to_nchw(x.double()), wr, stride=2).backward(nhwc_to_nchw(dy.double()))
    dw = torch.zeros((cin, 4, cout), dtype=torch.float32, device=DEV)
    dcat = torch.zeros((n, 2 * h, 2 * w, 2 * cout), dtype=dt, device=DEV)
    dcat[..., cout:] = dy.to(DEV)
    L.wgrad(x.to(DEV), dcat[..., cout:], dw, 2, algo=algo)
    assert relerr(dw, wr.grad.permute(0, 2, 3, 1).reshape(cin, 4, cout)) < 1e-4
    # 1x1 (first layer on its im2col operand)
    u = rnd((n, h, w, 64), dt, 15)
    s = rnd((n, h, w, 64), dt, 16)
    dw1 = torch.zeros((64, 1, 64), dtype=torch.float32, device=DEV)
    L.wgrad(u.to(DEV), s.to(DEV), dw1, 0, algo=algo)
    assert relerr(dw1.view(64, 64), u.double().reshape(-1, 64).t() @ s.double().reshape(-1, 64)) < 1e-4
    # channel sum (ConvTranspose bias gradient)
    out = torch.zeros(cout, dtype=torch.float32, device=DEV)
    L.channel_sum(dcat[..., cout:], out)
    assert relerr(out, dy.double().sum(dim=(0, 1, 2))) < 1e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("pool", [False, True], ids=["nopool", "pool"])
@pytest.mark.parametrize("n,h,w,c", [(2, 8, 8, 64), (3, 4, 6, 128), (1, 2, 2, 1024)])
def test_bn_relu_pool_forward_backward(dt, pool, n, h, w, c):
    g = torch.Generator().manual_seed(20)
    z = rnd((n, h, w, c), dt, 21, 2.0)
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g) * 0.3
    cbias = torch.randn(c, generator=g)
    rm0, rv0 = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5
    count = n * h * w
    # ---- reference (float64, NCHW, PyTorch ops the reference model uses) ----
    zr = nhwc_to_nchw(z.double()).requires_grad_(True)
    rm, rv = rm0.double().clone(), rv0.double().clone()
    y = F.batch_norm(zr + cbias.double()[None, :, None, None], rm, rv, gamma.double(), beta.double(), True, 0.1, 1e-5)
    a = torch.relu(y)
    # the CUDA path stores a in `dt`; pooling and ties are defined on the stored values
    a_q = a.detach().to(dt).double()
    outp = F.max_pool2d(a, 2) if pool else None
    # ---- CUDA ----
    zd = z.to(DEV)
    s1 = torch.zeros(c, dtype=torch.float64, device=DEV); s2 = torch.zeros_like(s1)
    L.bn_stats(zd, s1, s2)
    scale, shift, mean, invstd = (torch.empty(c, dtype=torch.float32, device=DEV) for _ in range(4))
    rmd, rvd = rm0.to(DEV), rv0.to(DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    L.bn_finalize(s1, s2, count, c, True, gamma.to(DEV), beta.to(DEV), cbias.to(DEV), rmd, rvd, nbt, 0.1, 1e-5,
                  scale, shift, mean, invstd)
    cat = torch.zeros((n, h, w, 2 * c), dtype=dt, device=DEV)
    ad = cat[..., :c]
    pooled = torch.empty((n, h // 2, w // 2, c), dtype=dt, device=DEV) if pool else None
    pidx = torch.empty((n, h // 2, w // 2, c // 8), dtype=torch.int16, device=DEV) if pool else None
    L.bn_relu_apply(zd, scale, shift, ad, pooled, pidx)
    assert relerr(ad, a.detach().permute(0, 2, 3, 1)) < TOL[dt]
    assert int(nbt) == 1
    assert relerr(rmd, rm) < 1e-5 and relerr(rvd, rv) < 1e-5
    if pool:
        assert torch.equal(pooled.cpu(), F.max_pool2d(nhwc_to_nchw(ad.cpu().float()), 2).permute(0, 2, 3, 1).to(dt))
    # ---- backward ----
    dy = rnd((n, h, w, c), dt, 22)
    dp = rnd((n, h // 2, w // 2, c), dt, 23) if pool else None
    if pool:
        # route the pooled gradient exactly like torch's max_pool2d backward does on the stored activations
        aq = nhwc_to_nchw(ad.cpu().double()).requires_grad_(True)
        F.max_pool2d(aq, 2).backward(nhwc_to_nchw(dp.double()))
        dA = nhwc_to_nchw(dy.double()) + aq.grad
    else:
        dA = nhwc_to_nchw(dy.double())
    a.backward(dA)
    sums = torch.zeros(2 * c, dtype=torch.float64, device=DEV)
    dz = torch.empty((n, h, w, c), dtype=dt, device=DEV)
    dgamma, dbeta = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    L.bn_relu_bwd(zd, dy.to(DEV), dp.to(DEV) if pool else None, scale, shift, mean, invstd, sums, dz, dgamma, dbeta,
                  pool_idx=pidx)
    ref_dz = zr.grad.permute(0, 2, 3, 1)
    # ReLU-mask decisions are taken on the stored (dt-rounded) activation: identical to the reference when a_q > 0 <=> a > 0
    assert torch.equal(a_q > 0, a.detach() > 0)
    assert relerr(dz, ref_dz) < TOL[dt]
    g_ref = torch.autograd.grad  # noqa: F841
    # dgamma/dbeta against closed form
    xhat = (zr.detach() - zr.detach().mean(dim=(0, 2, 3), keepdim=True)) / torch.sqrt(zr.detach().var(dim=(0, 2, 3), unbiased=False, keepdim=True) + 1e-5)
    dyy = dA * (a.detach() > 0)
    assert relerr(dbeta, dyy.sum(dim=(0, 2, 3))) < 1e-4
    assert relerr(dgamma, (dyy * xhat).sum(dim=(0, 2, 3))) < 2e-4


def test_bn_eval_mode_uses_running_stats():
    c = 64
    g = torch.Generator().manual_seed(30)
    gamma, beta, cbias = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g), torch.randn(c, generator=g)
    rm, rv = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5
    scale, shift = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    L.bn_finalize(None, None, 0, c, False, gamma.to(DEV), beta.to(DEV), cbias.to(DEV), rm.to(DEV), rv.to(DEV), None, 0.1, 1e-5,
                  scale, shift, None, None)
    z = torch.randn(4, c, generator=g)
    ref = (z + cbias - rm) / torch.sqrt(rv + 1e-5) * gamma + beta
    assert relerr(z.to(DEV) * scale + shift, ref) < 1e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("dout", [1, 3, 4])
def test_head_forward_backward(dt, dout):
    n, h, w, cin = 2, 16, 8, 64
    a = rnd((n, h, w, cin), dt, 40)
    wt = rnd((dout, cin, 1, 1), torch.float32, 41, 0.2)
    b = rnd((dout,), torch.float32, 42)
    ar = nhwc_to_nchw(a.double()).requires_grad_(True)
    wr, br = wt.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = F.conv2d(ar, wr, br)
    logits = torch.empty((n, dout, h, w), dtype=torch.float32, device=DEV)
    L.head_fprop(a.to(DEV), wt.to(DEV), b.to(DEV), dout, logits)
    assert relerr(logits, ref) < 1e-5
    dl = rnd((n, dout, h, w), torch.float32, 43)
    ref.backward(dl.double())
    da = torch.empty((n, h, w, cin), dtype=dt, device=DEV)
    dw = torch.zeros((dout, cin), dtype=torch.float32, device=DEV)
    db = torch.zeros(dout, dtype=torch.float32, device=DEV)
    L.head_bwd(dl.to(DEV), a.to(DEV), wt.to(DEV), dout, da, dw, db)
    assert relerr(da, ar.grad.permute(0, 2, 3, 1)) < TOL[dt]
    assert relerr(dw, wr.grad.view(dout, cin)) < 1e-5
    assert relerr(db, br.grad) < 1e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("dout,n,h,w,c", [(1, 2, 16, 8, 64), (3, 2, 12, 20, 64), (4, 3, 6, 10, 128), (3, 1, 64, 64, 64),
                                          # planes that are not multiples of the 16-pixel block / ragged last block /
                                          # more than one round of the persistent grid with blocks straddling two images
                                          (3, 2, 5, 7, 64), (2, 3, 37, 3, 64), (4, 1, 3, 3, 64), (3, 3, 150, 101, 64)])
def test_fused_head_and_batchnorm_backward(dt, dout, n, h, w, c):
    """unetk_head_bn_bwd_reduce/apply against autograd (float64) through  z -> BN(batch stats) -> ReLU -> 1x1 head."""
    z = rnd((n, h, w, c), dt, 60)
    gamma = rnd((c,), torch.float32, 61).abs() + 0.5
    beta = rnd((c,), torch.float32, 62, 0.3)
    wh = rnd((dout, c, 1, 1), torch.float32, 63, 0.2)
    bh = rnd((dout,), torch.float32, 64)
    dl = rnd((n, dout, h, w), torch.float32, 65)
    zr = nhwc_to_nchw(z.double()).requires_grad_(True)
    gr, br_ = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    whr, bhr = wh.double().requires_grad_(True), bh.double().requires_grad_(True)
    a = F.relu(F.batch_norm(zr, None, None, gr, br_, training=True, eps=1e-5))
    F.conv2d(a, whr, bhr).backward(dl.double())
    mean = z.double().mean((0, 1, 2))
    var = z.double().var((0, 1, 2), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    scale = (gamma.double() * invstd).float().to(DEV)
    shift = (beta.double() - mean * gamma.double() * invstd).float().to(DEV)
    sums = torch.zeros((3 + dout) * c, dtype=torch.float64, device=DEV)
    dz = torch.empty((n, h, w, c), dtype=dt, device=DEV)
    dgamma, dbeta = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    dwh, dbh = torch.empty(dout, c, device=DEV), torch.empty(dout, device=DEV)
    L.head_bn_bwd(dl.to(DEV), z.to(DEV), wh.to(DEV), dout, scale, shift, mean.float().to(DEV), invstd.float().to(DEV), sums,
                  dz, dgamma, dbeta, dwh, dbh)
    tol = TOL[dt]
    assert relerr(dz, zr.grad.permute(0, 2, 3, 1)) < tol
    assert relerr(dgamma, gr.grad) < tol and relerr(dbeta, br_.grad) < tol
    assert relerr(dwh, whr.grad.view(dout, c)) < tol
    assert relerr(dbh, bhr.grad) < 1e-5
    if c != 64:
        return
    # and against the unfused head kernel on the same inputs (the U-Net head has 64 input channels)
    a_dev = torch.empty((n, h, w, c), dtype=dt, device=DEV)
    L.bn_relu_apply(z.to(DEV), scale, shift, a_dev, None, None)
    da = torch.empty((n, h, w, c), dtype=dt, device=DEV)
    dw2, db2 = torch.zeros(dout, c, device=DEV), torch.zeros(dout, device=DEV)
    L.head_bwd(dl.to(DEV), a_dev, wh.to(DEV), dout, da, dw2, db2)
    assert relerr(dwh, dw2) < tol and relerr(dbh, db2) < 1e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("dout,n,h,w,c", [(1, 2, 16, 8, 64), (3, 2, 5, 7, 64), (4, 1, 3, 3, 64), (3, 1, 64, 64, 64), (2, 3, 37, 3, 64),
                                          (3, 3, 150, 101, 64)])
def test_fused_bn_relu_head_forward(dt, dout, n, h, w, c):
    """unetk_bn_relu_head_fprop == unetk_bn_relu_apply followed by unetk_head_fprop (same stored activation)."""
    z = rnd((n, h, w, c), dt, 70).to(DEV)
    scale = (rnd((c,), torch.float32, 71).abs() + 0.5).to(DEV)
    shift = rnd((c,), torch.float32, 72, 0.3).to(DEV)
    wh = rnd((dout, c, 1, 1), torch.float32, 73, 0.2).to(DEV)
    bh = rnd((dout,), torch.float32, 74).to(DEV)
    a_ref = torch.empty((n, h, w, c), dtype=dt, device=DEV)
    L.bn_relu_apply(z, scale, shift, a_ref, None, None)
    want = torch.empty((n, dout, h, w), dtype=torch.float32, device=DEV)
    L.head_fprop(a_ref, wh, bh, dout, want)
    for store in (False, True):
        a = torch.zeros((n, h, w, c), dtype=dt, device=DEV) if store else None
        got = torch.full((n, dout, h, w), float("nan"), dtype=torch.float32, device=DEV)
        L.bn_relu_head_fprop(z, scale, shift, a, wh, bh, dout, got)
        assert relerr(got, want) < 1e-5            # same products, different summation order
        if store:
            assert torch.equal(a, a_ref)
    ref = F.conv2d(nhwc_to_nchw(torch.relu(z.double().cpu() * scale.double().cpu() + shift.double().cpu())), wh.double().cpu(),
                   bh.double().cpu())
    assert relerr(want, ref) < TOL[dt]


def test_fused_head_rejects_more_than_four_classes():
    z = torch.zeros((1, 4, 4, 64), dtype=torch.float32, device=DEV)
    f = torch.zeros(64, device=DEV)
    with pytest.raises(RuntimeError, match="1..4 classes"):
        L.head_bn_bwd(torch.zeros((1, 5, 4, 4), device=DEV), z, torch.zeros((5, 64), device=DEV), 5, f, f, f, f,
                      torch.zeros(8 * 64, dtype=torch.float64, device=DEV), torch.empty_like(z), f.clone(), f.clone(),
                      torch.zeros((5, 64), device=DEV), torch.zeros(5, device=DEV))


def test_im2col_first_and_permute3():
    for n, cin, h, w in ((2, 3, 8, 8), (2, 3, 5, 150), (1, 4, 70, 67)):     # strips of 64 columns: ragged last strip
        x = rnd((n, cin, h, w), torch.float32, 50)
        for dt in (torch.float32, torch.bfloat16):
            out = torch.empty((n, h, w, 64), dtype=dt, device=DEV)
            L.im2col3x3_first(x.to(DEV), out)
            cols = F.unfold(x, 3, padding=1).view(n, cin, 9, h, w)          # [n][ci][t][h][w]
            ref = torch.zeros(n, h, w, 64)
            ref[..., :9 * cin] = cols.permute(0, 3, 4, 2, 1).reshape(n, h, w, 9 * cin)  # k = t*cin + ci
            assert torch.equal(out.cpu().float(), ref.to(dt).float())
    wt = rnd((8, 5, 9), torch.float32, 51)
    dst = torch.empty((5, 9, 8), dtype=torch.float32, device=DEV)
    # [co][ci][t] -> [ci][8-t][co]
    L.check(L.lib().unetk_permute3(wt.to(DEV).data_ptr(), dst.data_ptr() + 8 * 8 * 4, L.F32, 8, 5, 9, 45, 9, 1, 1, 72, -8,
                                   L.stream_ptr()))
    assert torch.equal(dst.cpu(), wt.flip(2).permute(1, 2, 0).contiguous())


def test_loss_against_golden_and_oracle(golden):
    from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss, WeightedMemoryEfficientDiceLoss
    g = golden["loss"]
    cases = json.loads(str(g["cases"]))
    for ci, case in enumerate(cases):
        kw = {k: v for k, v in case.items() if k != "c"}
        if "class_weights" in kw:
            kw["class_weights"] = torch.tensor(kw["class_weights"], dtype=torch.float32)
        logits = torch.from_numpy(g[f"logits_{ci}"]).to(DEV).requires_grad_(True)
        target = torch.from_numpy(g[f"target_{ci}"]).to(DEV)
        fn = WeightedDiceCELoss(**kw)
        loss = fn(logits, target)
        loss.backward()
        assert abs(loss.item() - float(g[f"loss_{ci}"])) < 2e-6, ci                 # vs the unmodified reference
        np.testing.assert_allclose(logits.grad.cpu().numpy(), g[f"grad_{ci}"], rtol=2e-4, atol=2e-8)
        ref = loss_oracle.dice_ce_loss(logits.detach().cpu(), target.cpu(), **kw)     # vs the float64 oracle
        assert abs(loss.item() - ref.item()) < 1e-6
        assert fn(logits.detach(), target.unsqueeze(1)).item() == loss.item()        # [N,1,H,W] form
        # scaled upstream gradient (gradient accumulation divides the loss, utils/training.py:49)
        lg2 = logits.detach().clone().requires_grad_(True)
        (fn(lg2, target) / 4).backward()
        np.testing.assert_allclose(lg2.grad.cpu().numpy(), logits.grad.cpu().numpy() / 4, rtol=1e-6, atol=1e-12)
        with pytest.raises(ValueError):
            fn(logits.detach(), target.unsqueeze(1).repeat(1, 2, 1, 1))
    # standalone Dice
    logits = torch.from_numpy(g["logits_0"]).to(DEV)
    target = torch.from_numpy(g["target_0"]).to(DEV)
    d = WeightedMemoryEfficientDiceLoss(smooth=1.0)(logits, target.unsqueeze(1))
    ref = loss_oracle.dice_ce_loss(logits.cpu(), target.cpu(), smooth_dice=1.0, ce_weight=0.0)
    assert abs(d.item() - ref.item()) < 1e-6


def test_loss_large_random_vs_oracle():
    from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss
    g = torch.Generator().manual_seed(60)
    logits = torch.randn(4, 4, 64, 96, generator=g) * 3
    target = torch.randint(0, 4, (4, 64, 96), generator=g)
    w = torch.tensor([0.2, 1.0, 1.2, 1.5])
    fn = WeightedDiceCELoss(smooth_dice=1e-5, class_weights=w, ignore_index=3)
    lg = logits.to(DEV).requires_grad_(True)
    loss = fn(lg, target.to(DEV))
    loss.backward()
    ref = loss_oracle.dice_ce_loss(logits, target, smooth_dice=1e-5, class_weights=w, ignore_index=3)
    gref = loss_oracle.dice_ce_grad(logits, target, smooth_dice=1e-5, class_weights=w, ignore_index=3)
    assert abs(loss.item() - ref.item()) < 1e-6 * max(1, abs(ref.item()))
    assert relerr(lg.grad, gref) < 1e-5


def test_loss_out_of_range_label_raises():
    from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss
    import os
    os.environ["UNETK_STRICT_LABELS"] = "1"
    try:
        fn = WeightedDiceCELoss()
        with pytest.raises(RuntimeError):
            fn(torch.zeros(1, 3, 4, 4, device=DEV), torch.full((1, 4, 4), 3, device=DEV))
    finally:
        os.environ.pop("UNETK_STRICT_LABELS")


def test_metrics_bit_exact_against_golden_and_oracle(golden):
    from image_segmentation_b200.utils.MetricsHistory import MetricsHistory
    g = golden["metrics"]
    cases = json.loads(str(g["cases"]))
    for ci, case in enumerate(cases):
        c, ign = case["c"], case["ignore_index"]
        agg = MetricsHistory(c, ign)
        for pred, label in zip(g[f"pred_{ci}"], g[f"label_{ci}"]):
            agg.accumulate(torch.from_numpy(pred).to(DEV), torch.from_numpy(label).to(DEV))
        counts = np.stack([agg.total_tp.numpy(), agg.total_fp.numpy(), agg.total_fn.numpy(), agg.total_tn.numpy()])
        np.testing.assert_array_equal(counts, g[f"counts_{ci}"])                   # bit-exact vs the reference
        md, mi, ma = agg.compute_epoch_metrics()
        np.testing.assert_allclose([md, mi, ma], g[f"means_{ci}"], rtol=1e-12)
        agg.reset()
        assert float(agg.total_tp.sum()) == 0.0
    # batch form + NaN/tie semantics vs the numpy oracle
    gen = torch.Generator().manual_seed(70)
    pred = torch.randn(5, 4, 33, 17, generator=gen)
    pred[:, :, ::2, ::3] = pred[:, :1, ::2, ::3]
    pred[0, 2, 0, 0] = float("nan")
    label = torch.randint(0, 4, (5, 33, 17), generator=gen)
    agg = MetricsHistory(4)
    agg.accumulate(pred.to(DEV), label.to(DEV))
    tot = np.zeros((4, 4), dtype=np.int64)
    for p, l in zip(pred.numpy(), label.numpy()):
        tot += np.stack(metrics_oracle.confusion_counts(p, l, 4))
    got = np.stack([agg.total_tp.numpy(), agg.total_fp.numpy(), agg.total_fn.numpy(), agg.total_tn.numpy()]).astype(np.int64)
    np.testing.assert_array_equal(got, tot)
    import pickle
    agg2 = pickle.loads(pickle.dumps(agg))
    assert torch.equal(agg2.total_tp, agg.total_tp)
    bad = MetricsHistory(3)
    bad.accumulate(torch.zeros(3, 2, 2, device=DEV), torch.full((2, 2), 3, device=DEV))
    with pytest.raises(RuntimeError):
        bad.total_tp


def test_cpu_tensors_are_rejected():
    from image_segmentation_b200.unet.unet import unet
    from image_segmentation_b200.utils.weighted_loss import WeightedDiceCELoss
    with pytest.raises(RuntimeError):
        unet(3, 3)(torch.zeros(1, 3, 16, 16))
    with pytest.raises(RuntimeError):
        WeightedDiceCELoss()(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))


def test_batched_weight_pack_and_unpack():
    g = torch.Generator().manual_seed(80)
    w3 = torch.randn(64, 96, 3, 3, generator=g)        # conv3x3  [co][ci][3][3]
    wT = torch.randn(128, 64, 2, 2, generator=g)       # convT    [ci][co][2][2]
    w1 = torch.randn(64, 3, 3, 3, generator=g)         # first conv
    for dt in (torch.float32, torch.bfloat16):
        d3, dT, d1 = w3.to(DEV), wT.to(DEV), w1.to(DEV)
        wf3 = torch.empty((64, 9, 96), dtype=dt, device=DEV); wd3 = torch.empty((96, 9, 64), dtype=dt, device=DEV)
        wfT = torch.empty((4 * 64, 128), dtype=dt, device=DEV); wdT = torch.empty((128, 4, 64), dtype=dt, device=DEV)
        wf1 = torch.zeros((64, 64), dtype=dt, device=DEV)
        jobs = L.WeightJobs([(d3.data_ptr(), wf3.data_ptr(), wd3.data_ptr(), 0, 64, 96, 0),
                             (dT.data_ptr(), wfT.data_ptr(), wdT.data_ptr(), 1, 64, 128, 0),
                             (d1.data_ptr(), wf1.data_ptr(), None, 2, 64, 3, 64)], DEV)
        L.weights_pack(jobs, dt)
        assert torch.equal(wf3.cpu(), w3.permute(0, 2, 3, 1).reshape(64, 9, 96).to(dt))
        assert torch.equal(wd3.cpu(), w3.flip(2, 3).permute(1, 2, 3, 0).reshape(96, 9, 64).to(dt))
        assert torch.equal(wfT.cpu(), wT.permute(2, 3, 1, 0).reshape(256, 128).to(dt))
        assert torch.equal(wdT.cpu(), wT.permute(0, 2, 3, 1).reshape(128, 4, 64).to(dt))
        ref1 = torch.zeros(64, 64)
        ref1[:, :27] = w1.permute(0, 2, 3, 1).reshape(64, 27)
        assert torch.equal(wf1.cpu(), ref1.to(dt))
    # unpack: gradient workspaces (operand layout) -> parameter layout, destinations as offsets into one flat buffer
    ws3 = torch.randn(64, 9, 96, generator=g).to(DEV)
    wsT = torch.randn(128, 4, 64, generator=g).to(DEV)
    ws1 = torch.randn(64, 64, generator=g).to(DEV)
    flat = torch.zeros(7 + w3.numel() + wT.numel() + w1.numel(), device=DEV)
    o3, oT, o1 = 7, 7 + w3.numel(), 7 + w3.numel() + wT.numel()
    jobs = L.WeightJobs([(ws3.data_ptr(), 4 * o3, None, 0, 64, 96, 0), (wsT.data_ptr(), 4 * oT, None, 1, 64, 128, 0),
                         (ws1.data_ptr(), 4 * o1, None, 2, 64, 3, 64)], DEV)
    L.weights_unpack(jobs, flat)
    assert torch.equal(flat[o3:oT].view(64, 96, 3, 3).cpu(), ws3.cpu().view(64, 3, 3, 96).permute(0, 3, 1, 2))
    assert torch.equal(flat[oT:o1].view(128, 64, 2, 2).cpu(), wsT.cpu().view(128, 2, 2, 64).permute(0, 3, 1, 2))
    assert torch.equal(flat[o1:].view(64, 3, 3, 3).cpu(), ws1.cpu()[:, :27].view(64, 3, 3, 3).permute(0, 3, 1, 2))
    assert float(flat[:7].abs().sum()) == 0.0


@pytest.mark.parametrize("dt,algo", ALGOS, ids=IDS)
@pytest.mark.parametrize("mode,n,h,w,cin,cout", [("3x3", 2, 32, 16, 64, 64), ("3x3", 1, 16, 16, 128, 256), ("3x3", 3, 8, 8, 64, 128),
                                                   ("3x3", 2, 48, 24, 128, 128), ("gather", 2, 8, 8, 64, 128)])
def test_data_gradient_with_fused_bn_backward_reduction(dt, algo, mode, n, h, w, cin, cout):
    """unetk_conv(bn_reduce=...) must produce the same sums as the stand-alone unetk_bn_relu_bwd_reduce on its output."""
    g = torch.Generator().manual_seed(90)
    if mode == "3x3":
        x = rnd((n, h, w, cin), dt, 91)
        wt = rnd((cout, 9, cin), dt, 92, 0.05)
        y = torch.empty((n, h, w, cout), dtype=dt, device=DEV)
        m = L.MODE_3X3
    else:
        x = rnd((n, 2 * h, 2 * w, cin), dt, 91)
        wt = rnd((cout, 4, cin), dt, 92, 0.05)
        y = torch.empty((n, h, w, cout), dtype=dt, device=DEV)
        m = L.MODE_CONVT_GATHER
    z = rnd((n, h, w, cout), dt, 93, 2.0).to(DEV)
    scale = (torch.rand(cout, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(cout, generator=g) * 0.5).to(DEV)
    mean = (torch.randn(cout, generator=g) * 0.3).to(DEV)
    invstd = (torch.rand(cout, generator=g) + 0.5).to(DEV)
    fused = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    L.conv(x.to(DEV), wt.to(DEV), y, m, algo=algo, bn_reduce=(z, scale, shift, mean, invstd, fused))
    ref = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    args = L.BnBwdArgs(L.nhwc(z), L.nhwc(y), L.nhwc(None), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                       ref.data_ptr(), L.nhwc(None), None, None, None)
    L.check(L.lib().unetk_bn_relu_bwd_reduce(__import__("ctypes").byref(args), L.stream_ptr()))
    torch.cuda.synchronize()
    scale_ref = ref.abs().max().item()
    assert (fused - ref).abs().max().item() < 2e-4 * scale_ref, ((fused - ref).abs().max().item(), scale_ref)
    # and the data gradient itself is unchanged by the fusion
    y2 = torch.empty_like(y)
    L.conv(x.to(DEV), wt.to(DEV), y2, m, algo=algo)
    assert torch.equal(y, y2)


def _guarded(shape, dtype, fill=0.0):
    """A tensor view with 4096 sentinel elements on either side (out-of-bounds writes become visible)."""
    n = int(np.prod(shape))
    base = torch.full((n + 8192,), 12345.0, dtype=dtype, device=DEV)
    view = base[4096:4096 + n].view(*shape)
    view.fill_(fill)
    return base, view


def _guards_intact(base, n):
    return bool((base[:4096] == 12345.0).all()) and bool((base[4096 + n:] == 12345.0).all())


def test_new_kernels_do_not_write_outside_their_outputs():
    """Sentinel regions around every output of the kernels added late in round 1 (compute-sanitizer is not available on
    the GPU pool): tiled im2col, nine-tap 64-channel wgrad, fused head forward / backward, ragged crop-resize."""
    # im2col: ragged strip and row remainders
    n, cin, h, w = 2, 3, 19, 83
    x = rnd((n, cin, h, w), torch.float32, 80).to(DEV)
    base, out = _guarded((n, h, w, 64), torch.bfloat16)
    L.im2col3x3_first(x, out)
    assert _guards_intact(base, out.numel())
    # 64-channel 3x3 weight gradient with partial tiles (h, w not multiples of 16 / 8)
    n, h, w = 2, 40, 20
    dy = rnd((n, h, w, 64), torch.bfloat16, 81).to(DEV)
    xa = rnd((n, h, w, 128), torch.bfloat16, 82).to(DEV)
    base, dw = _guarded((64, 9, 128), torch.float32)
    L.wgrad(dy, xa, dw, 1, algo=L.ALGO_TC)
    assert _guards_intact(base, dw.numel())
    ref = torch.zeros(64, 9, 128, dtype=torch.float64)
    xp = F.pad(nhwc_to_nchw(xa.double().cpu()), (1, 1, 1, 1))
    dyc = nhwc_to_nchw(dy.double().cpu())
    for t in range(9):
        r, s_ = t // 3, t % 3
        ref[:, t, :] = torch.einsum("nchw,ndhw->cd", dyc, xp[:, :, r:r + h, s_:s_ + w])
    assert relerr(dw, ref) < 1e-4
    # fused head forward / backward
    n, h, w, c, dout = 2, 9, 11, 64, 3
    z = rnd((n, h, w, c), torch.bfloat16, 83).to(DEV)
    vec = [(rnd((c,), torch.float32, 84 + i).abs() + 0.5).to(DEV) for i in range(4)]
    wh, bh = rnd((dout, c), torch.float32, 90, 0.2).to(DEV), rnd((dout,), torch.float32, 91).to(DEV)
    base, logits = _guarded((n, dout, h, w), torch.float32)
    L.bn_relu_head_fprop(z, vec[0], vec[1], None, wh, bh, dout, logits)
    assert _guards_intact(base, logits.numel())
    base, dz = _guarded((n, h, w, c), torch.bfloat16)
    sums = torch.zeros((3 + dout) * c, dtype=torch.float64, device=DEV)
    dl = rnd((n, dout, h, w), torch.float32, 92).to(DEV)
    outs = [torch.empty(c, device=DEV), torch.empty(c, device=DEV), torch.empty(dout, c, device=DEV), torch.empty(dout, device=DEV)]
    L.head_bn_bwd(dl, z, wh, dout, vec[0], vec[1], vec[2], vec[3], sums, dz, *outs)
    assert _guards_intact(base, dz.numel())
    # ragged crop-resize
    from image_segmentation_b200.utils.utils import process_batch_reverse
    metas = [{"original_size": (33, 21), "new_size": (16, 10), "pad": (3, 0, 3, 0), "scale": 0.5},
             {"original_size": (7, 50), "new_size": (2, 16), "pad": (0, 7, 0, 7), "scale": 0.3}]
    res = process_batch_reverse(rnd((2, 4, 16, 16), torch.float32, 93).to(DEV), metas)
    assert [tuple(r.shape) for r in res] == [(4, 33, 21), (4, 7, 50)] and all(torch.isfinite(r).all() for r in res)


@pytest.mark.parametrize("mode,n,h,w,cu,cs", [(1, 8, 64, 64, 64, 64), (1, 4, 32, 32, 128, 256), (2, 8, 16, 16, 128, 64), (0, 2, 64, 64, 64, 64)])
def test_deterministic_wgrad_matches_atomic_wgrad_and_is_reproducible(mode, n, h, w, cu, cs):
    """UNETK_TC_DETERMINISTIC: same result as the atomic reduction (fp32 rounding of a different summation order), identical
    bits across repeated launches, accumulation semantics preserved, scratch size from unetk_wgrad_partial_bytes."""
    dt = torch.bfloat16
    taps = (1, 9, 4)[mode]
    u = rnd((n, h, w, cu), dt, 70).to(DEV)
    s = rnd((n, 2 * h, 2 * w, cs) if mode == 2 else (n, h, w, cs), dt, 71).to(DEV)
    need = L.wgrad_partial_bytes(u, s, mode, L.ALGO_TC)
    ref = torch.zeros((cu, taps, cs), dtype=torch.float32, device=DEV)
    L.wgrad(u, s, ref, mode, algo=L.ALGO_TC)
    if need == 0:
        pytest.skip("single split on this device: nothing to reduce")
    partial = torch.empty(need // 4, dtype=torch.float32, device=DEV)
    outs = []
    for _ in range(3):
        dw = torch.zeros_like(ref)
        L.wgrad(u, s, dw, mode, algo=L.ALGO_TC, partial=partial)
        outs.append(dw)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert relerr(outs[0], ref) < 1e-5
    L.wgrad(u, s, outs[0], mode, algo=L.ALGO_TC, partial=partial)          # a second call accumulates
    assert relerr(outs[0], 2 * ref) < 1e-5
    with pytest.raises(RuntimeError, match="partial"):
        L.wgrad(u, s, dw, mode, algo=L.ALGO_TC, partial=partial[:16])
