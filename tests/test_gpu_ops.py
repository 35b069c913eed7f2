"""-m gpu: per-op parity of every kernel behind the C ABI against the matching torch op on CPU (the oracle's ops),
on identical inputs.  Integer results (pool indices, argmax masks) must be bit-exact; floating point within the
tolerance written at each assert (fp32 storage: 1e-4..1e-3 rel; bf16 storage: 1e-2 rel, one bf16 ulp = 2^-8)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import *  # noqa: E402,F401,F403

DEFAULT_APPLY_STREAM = 8      # library default of tuning key 3 (restored after the A/B test)

DTS = [F32, BF16]
TOL = {F32: 2e-4, BF16: 1.2e-2}


@pytest.fixture(autouse=True)
def _need_gpu(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    call("unetca_set_conv_impl", 0)
    yield
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,C,H,W", [(2, 64, 8, 16), (1, 192, 6, 4), (3, 1024, 2, 2), (2, 64, 5, 13), (1, 128, 7, 6)])
def test_maxpool_bit_exact(dt, B, C, H, W):
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.randint(0, 3, (B, C, H, W)).astype(np.float32))      # many exact ties
    x[0, 0, 0, 1] = float("nan")
    x[-1, C - 1, H - 1, W - 2] = float("nan")
    xd = to_nhwc(x, dt)
    pooled = torch.empty(B, H // 2, W // 2, C, dtype=TDT[dt], device="cuda")
    pos = torch.empty(B, H // 2, W // 2, C, dtype=torch.uint8, device="cuda")
    idx = torch.empty(B, C, H // 2, W // 2, dtype=torch.int64, device="cuda")
    call("unetca_maxpool2x2", dt, ptr(xd), C, ptr(pooled), C, ptr(pos), ptr(idx), B, H, W, C, stream())
    ry, ridx = F.max_pool2d(x, 2, return_indices=True)
    assert torch.equal(idx.cpu(), ridx)                                          # bit-exact indices
    got = from_nhwc(pooled)
    assert torch.equal(torch.isnan(got), torch.isnan(ry))
    assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(ry))
    # backward: unpool + skip add
    dpool = torch.from_numpy(rs.standard_normal((B, C, H // 2, W // 2)).astype(np.float32))
    skip = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    dx = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    dpd, skd = to_nhwc(dpool, dt), to_nhwc(skip, dt)
    call("unetca_pool_bwd_add", dt, ptr(skd), C, ptr(dpd), C, ptr(pos), ptr(dx), C, B, H, W, C, stream())
    ref = rounded(skip, dt) + F.max_unpool2d(rounded(dpool, dt), ridx, 2, output_size=(H, W))
    assert relerr(from_nhwc(dx), ref) < TOL[dt]


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,C,h,w,H,W", [(2, 64, 4, 6, 5, 6), (1, 128, 10, 12, 10, 13), (2, 64, 24, 12, 25, 13)])
def test_resize_guard_bilinear(dt, B, C, h, w, H, W):
    """Decoder resize guard (UCA:138-157): F_T.resize(tensor, BILINEAR) == interpolate(antialias=True), and its adjoint."""
    rs = np.random.RandomState(9)
    x = torch.from_numpy(rs.standard_normal((B, C, h, w)).astype(np.float32))
    xr = rounded(x, dt).requires_grad_(True)
    ref = F.interpolate(xr, size=(H, W), mode="bilinear", align_corners=False, antialias=True)
    dy = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    ref.backward(rounded(dy, dt))
    xd = to_nhwc(x, dt)
    cat = torch.zeros(B, H, W, 2 * C, dtype=TDT[dt], device="cuda")             # written into the upper channel half
    call("unetca_resize_bilinear_fwd", dt, ptr(xd), C, h, w, ptr(cat[..., C:]), 2 * C, H, W, B, C, stream())
    assert relerr(from_nhwc(cat[..., C:]), ref.detach()) < TOL[dt]
    assert cat[..., :C].abs().max().item() == 0
    dcat = torch.zeros(B, H, W, 2 * C, dtype=TDT[dt], device="cuda")
    dcat[..., C:] = to_nhwc(dy, dt)
    dx = torch.empty(B, h, w, C, dtype=TDT[dt], device="cuda")
    call("unetca_resize_bilinear_bwd", dt, ptr(dcat[..., C:]), 2 * C, H, W, ptr(dx), C, h, w, B, C, stream())
    assert relerr(from_nhwc(dx), xr.grad) < TOL[dt]


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,C,H,W", [(2, 64, 16, 16), (3, 128, 8, 4)])
def test_bn_relu_train_and_backward(dt, B, C, H, W):
    rs = np.random.RandomState(1)
    y = torch.from_numpy((rs.standard_normal((B, C, H, W)) * 2 + 0.5).astype(np.float32))
    yr = rounded(y, dt)
    gamma = torch.from_numpy((1 + 0.1 * rs.standard_normal(C)).astype(np.float32))
    beta = torch.from_numpy((0.1 * rs.standard_normal(C)).astype(np.float32))
    cbias = torch.from_numpy((0.3 * rs.standard_normal(C)).astype(np.float32))
    rm, rv = torch.zeros(C), torch.ones(C)
    yd = to_nhwc(y, dt)
    parts = parts_buf(B)
    n = cint()
    call("unetca_chan_stats", dt, ptr(yd), C, C, B * H * W, ptr(parts), ctypes.byref(n), stream())
    g = {k: v.cuda() for k, v in dict(gamma=gamma, beta=beta, cbias=cbias, rm=rm.clone(), rv=rv.clone()).items()}
    mean, invstd, scale, shift = (torch.empty(C, device="cuda") for _ in range(4))
    call("unetca_bn_finalize_train", ptr(parts), n.value, C, B * H * W, ptr(g["cbias"]), ptr(g["gamma"]), ptr(g["beta"]),
         ptr(g["rm"]), ptr(g["rv"]), 0.1, 1e-5, ptr(mean), ptr(invstd), ptr(scale), ptr(shift), stream())
    out = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    call("unetca_bn_relu", dt, ptr(yd), C, ptr(out), C, B, H * W, C, ptr(scale), ptr(shift), None, None, stream())
    # reference: conv output = yr + bias, BN train, ReLU
    yin = (yr + cbias[None, :, None, None]).requires_grad_(True)
    rrm, rrv = rm.clone(), rv.clone()
    ref = F.relu(F.batch_norm(yin, rrm, rrv, gamma, beta, True, 0.1, 1e-5))
    assert relerr(from_nhwc(out), ref.detach()) < TOL[dt]
    assert relerr(g["rm"].cpu(), rrm) < 1e-5 and relerr(g["rv"].cpu(), rrv) < 1e-5
    # backward
    dout = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    dor = rounded(dout, dt)
    gw = torch.autograd.grad(ref, [yin], dor, retain_graph=True)[0]
    gam = gamma.clone().requires_grad_(True); bet = beta.clone().requires_grad_(True)
    ref2 = F.relu(F.batch_norm(yin.detach(), None, None, gam, bet, True, 0.1, 1e-5))
    dgam, dbet = torch.autograd.grad(ref2, [gam, bet], dor)
    dd = to_nhwc(dout, dt)
    call("unetca_bn_bwd_reduce", dt, ptr(dd), C, ptr(yd), C, B, H * W, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
         None, None, ptr(parts), ctypes.byref(n), stream())
    dgamma, dbeta = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    coef = torch.empty(3, C, device="cuda")
    call("unetca_bn_bwd_finalize", ptr(parts), n.value, C, B * H * W, ptr(g["gamma"]), ptr(invstd), ptr(dgamma), ptr(dbeta),
         ptr(coef), stream())
    dy = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    call("unetca_bn_bwd_apply", dt, ptr(dd), C, ptr(yd), C, ptr(dy), C, B, H * W, C, ptr(scale), ptr(shift), ptr(mean),
         ptr(invstd), None, None, ptr(coef), stream())
    assert relerr(dgamma.cpu(), dgam) < 1e-3 and relerr(dbeta.cpu(), dbet) < 1e-3
    assert relerr(from_nhwc(dy), gw) < TOL[dt]


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,C,H,W,pool", [(2, 64, 16, 16, True), (3, 128, 4, 8, False)])
def test_se_block_forward_backward(dt, B, C, H, W, pool):
    """relu(bn) -> SE (squeeze, FC, sigmoid, scale) [-> maxpool] and its backward, fused kernels vs torch ops."""
    rs = np.random.RandomState(2)
    Cr = C // 16
    y = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    yr = rounded(y, dt)
    scale = torch.from_numpy((1 + 0.2 * rs.standard_normal(C)).astype(np.float32))
    shift = torch.from_numpy((0.2 * rs.standard_normal(C)).astype(np.float32))
    w1 = torch.from_numpy((rs.standard_normal((Cr, C)) / np.sqrt(C)).astype(np.float32))
    w2 = torch.from_numpy((rs.standard_normal((C, Cr)) / np.sqrt(Cr)).astype(np.float32))
    yd = to_nhwc(y, dt)
    sc, sh, w1d, w2d = scale.cuda(), shift.cuda(), w1.cuda(), w2.cuda()
    parts = parts_buf(B)
    n = cint()
    call("unetca_bn_relu", dt, ptr(yd), C, None, 0, B, H * W, C, ptr(sc), ptr(sh), ptr(parts), ctypes.byref(n), stream())
    p, z, s = torch.empty(B, C, device="cuda"), torch.empty(B, Cr, device="cuda"), torch.empty(B, C, device="cuda")
    call("unetca_se_fc", ptr(parts), n.value, B, C, Cr, H * W, ptr(w1d), ptr(w2d), ptr(p), ptr(z), ptr(s), stream())
    # output into the lower channel half of a 2C-wide concat buffer (skip-first torch.cat, UCA:140)
    cat = torch.zeros(B, H, W, 2 * C, dtype=TDT[dt], device="cuda")
    pooled = torch.empty(B, H // 2, W // 2, C, dtype=TDT[dt], device="cuda") if pool else None
    pos = torch.empty(B, H // 2, W // 2, C, dtype=torch.uint8, device="cuda") if pool else None
    call("unetca_se_scale_pool", dt, ptr(yd), C, ptr(cat), 2 * C, ptr(pooled), C if pool else 0, ptr(pos), B, H, W, C,
         ptr(sc), ptr(sh), ptr(s), stream())
    # reference
    a = F.relu(yr * scale[None, :, None, None] + shift[None, :, None, None]).requires_grad_(True)
    w1r, w2r = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    sr = torch.sigmoid(F.linear(F.relu(F.linear(a.mean((2, 3)), w1r)), w2r))
    o = a * sr[:, :, None, None]
    assert relerr(s.cpu(), sr.detach()) < 1e-4
    got = from_nhwc(cat[..., :C])
    assert relerr(got, o.detach()) < TOL[dt]
    assert cat[..., C:].abs().max().item() == 0
    if pool:
        ry, ridx = F.max_pool2d(got, 2, return_indices=True)                     # pool of the *stored* values
        assert torch.equal(from_nhwc(pooled), ry)
        code = pos.cpu().permute(0, 3, 1, 2).long()
        ho = torch.arange(H // 2)[None, None, :, None]; wo = torch.arange(W // 2)[None, None, None, :]
        assert torch.equal((2 * ho + code // 2) * W + 2 * wo + code % 2, ridx)   # bit-exact window positions
    # backward
    dout = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    dor = rounded(dout, dt)
    da, dw1, dw2 = torch.autograd.grad(o, [a, w1r, w2r], dor)
    dz_ref = da * (a > 0)
    dd = to_nhwc(dout, dt)
    call("unetca_se_bwd_reduce", dt, ptr(dd), C, ptr(yd), C, B, H * W, C, ptr(sc), ptr(sh), ptr(parts), ctypes.byref(n),
         stream())
    dpre2, dzz, dp = torch.empty(B, C, device="cuda"), torch.empty(B, Cr, device="cuda"), torch.empty(B, C, device="cuda")
    gw1, gw2 = torch.empty(Cr, C, device="cuda"), torch.empty(C, Cr, device="cuda")
    call("unetca_se_fc_bwd", ptr(parts), n.value, B, C, Cr, ptr(w1d), ptr(w2d), ptr(p), ptr(z), ptr(s), ptr(dpre2), ptr(dzz),
         ptr(dp), ptr(gw1), ptr(gw2), stream())
    assert relerr(gw1.cpu(), dw1) < 2e-3 and relerr(gw2.cpu(), dw2) < 2e-3
    # the SE-scale / squeeze derivative folded into the BN backward: with mean=0, invstd=1 the "xhat" sums are
    # just sum(dz*y); check sum(dz) per channel and dy through a BN with gamma*invstd = 1, c1 = c2 = 0
    zeros, ones = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    call("unetca_bn_bwd_reduce", dt, ptr(dd), C, ptr(yd), C, B, H * W, C, ptr(sc), ptr(sh), ptr(zeros), ptr(ones), ptr(s),
         ptr(dp), ptr(parts), ctypes.byref(n), stream())
    sums = parts[: n.value * 2 * C].view(n.value, 2, C).sum(0).cpu()
    assert relerr(sums[0], dz_ref.sum((0, 2, 3))) < 2e-3
    assert relerr(sums[1], (dz_ref * yr).sum((0, 2, 3))) < 2e-3
    coef = torch.stack([ones, zeros, zeros]).contiguous()
    dy = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    call("unetca_bn_bwd_apply", dt, ptr(dd), C, ptr(yd), C, ptr(dy), C, B, H * W, C, ptr(sc), ptr(sh), ptr(zeros), ptr(ones),
         ptr(s), ptr(dp), ptr(coef), stream())
    assert relerr(from_nhwc(dy), dz_ref) < TOL[dt]
    # merged SE + BN reduction (one pass over dO, Y2): same FC-chain outputs and the same BN finalize results as the
    # two-pass path, with a non-trivial mean / invstd
    mean = torch.from_numpy((0.3 * rs.standard_normal(C)).astype(np.float32)).cuda()
    invstd = torch.from_numpy((1 + 0.1 * rs.random_sample(C)).astype(np.float32)).cuda()
    gamma = torch.from_numpy((1 + 0.2 * rs.standard_normal(C)).astype(np.float32)).cuda()
    call("unetca_bn_bwd_reduce", dt, ptr(dd), C, ptr(yd), C, B, H * W, C, ptr(sc), ptr(sh), ptr(mean), ptr(invstd), ptr(s),
         ptr(dp), ptr(parts), ctypes.byref(n), stream())
    dg_a, db_a, coef_a = torch.empty(C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(3, C, device="cuda")
    call("unetca_bn_bwd_finalize", ptr(parts), n.value, C, B * H * W, ptr(gamma), ptr(invstd), ptr(dg_a), ptr(db_a),
         ptr(coef_a), stream())
    parts4 = parts_buf(B, 4096)
    call("unetca_se_squeeze", dt, ptr(yd), C, B, H * W, C, ptr(sc), ptr(sh), ptr(parts4), ctypes.byref(n), stream())
    p3, z3, s3 = torch.empty_like(p), torch.empty_like(z), torch.empty_like(s)
    sums34 = torch.empty(B, 2, C, device="cuda")
    call("unetca_se_fc3", ptr(parts4), n.value, B, C, Cr, H * W, ptr(w1d), ptr(w2d), ptr(sc), ptr(sh), ptr(mean), ptr(p3),
         ptr(z3), ptr(s3), ptr(sums34), stream())
    assert relerr(p3, p) < 1e-5 and relerr(z3, z) < 1e-4 and relerr(s3, s) < 1e-5
    on = (yr * scale[None, :, None, None] + shift[None, :, None, None]) > 0
    assert torch.equal(sums34[:, 0].cpu(), on.float().sum((2, 3)))
    assert relerr(sums34[:, 1].cpu(), ((yr - mean.cpu()[None, :, None, None]) * on).sum((2, 3))) < 1e-4
    call("unetca_se_bn_bwd_reduce", dt, ptr(dd), C, ptr(yd), C, B, H * W, C, ptr(sc), ptr(sh), ptr(mean), ptr(parts4),
         ctypes.byref(n), stream())
    sums4 = torch.empty(B, 4, C, device="cuda")
    dpre2b, dzb, dpb = torch.empty_like(dpre2), torch.empty_like(dzz), torch.empty_like(dp)
    gw1b, gw2b = torch.empty_like(gw1), torch.empty_like(gw2)
    call("unetca_se_fc_bwd_fused", ptr(parts4), n.value, B, C, Cr, ptr(w1d), ptr(w2d), ptr(p), ptr(z), ptr(s), ptr(sc),
         ptr(sh), ptr(mean), ptr(sums34), ptr(sums4), ptr(dpre2b), ptr(dzb), ptr(dpb), ptr(gw1b), ptr(gw2b), stream())
    assert relerr(dpb, dp) < 1e-4 and relerr(dpre2b, dpre2) < 1e-4
    assert relerr(gw1b, gw1) < 1e-4 and relerr(gw2b, gw2) < 1e-4
    dg_b, db_b, coef_b = torch.empty(C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(3, C, device="cuda")
    call("unetca_bn_bwd_finalize_se", ptr(sums4), B, C, B * H * W, H * W, ptr(gamma), ptr(invstd), ptr(s), ptr(dpb),
         ptr(dg_b), ptr(db_b), ptr(coef_b), stream())
    assert relerr(dg_b, dg_a) < 1e-4 and relerr(db_b, db_a) < 1e-4 and relerr(coef_b, coef_a) < 1e-4


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("nc", [2, 3])
def test_outc_and_cross_entropy(dt, nc):
    rs = np.random.RandomState(3)
    B, C, H, W = 2, 64, 16, 24
    x = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    xr = rounded(x, dt).requires_grad_(True)
    w = torch.from_numpy((rs.standard_normal((nc, C, 1, 1)) / 8).astype(np.float32)).requires_grad_(True)
    b = torch.from_numpy((0.1 * rs.standard_normal(nc)).astype(np.float32)).requires_grad_(True)
    t = torch.from_numpy(rs.randint(0, nc, (B, H, W)).astype(np.int64))
    t[0, :2, :] = 255
    xd = to_nhwc(x, dt)
    wd, bd, td = w.detach().cuda(), b.detach().cuda(), t.cuda()
    logits = torch.empty(B, nc, H, W, device="cuda")
    call("unetca_outc_fwd", dt, ptr(xd), C, C, ptr(wd), ptr(bd), nc, ptr(logits), B, H * W, stream())
    ref = F.conv2d(xr, w, b)
    assert relerr(logits.cpu(), ref.detach()) < 1e-5
    loss_ref = F.cross_entropy(ref, t, ignore_index=255)
    loss_ref.backward()
    g = torch.empty_like(logits)
    mask = torch.empty(B, H, W, dtype=torch.int64, device="cuda")
    parts = parts_buf(B)
    out, gs = torch.empty(2, device="cuda"), torch.empty(1, device="cuda")
    call("unetca_cross_entropy", ptr(logits), ptr(td), nc, B, H * W, 255, None, ptr(g), ptr(mask), ptr(parts), ptr(out),
         ptr(gs), stream())
    assert abs(out[0].item() - loss_ref.item()) < 1e-5 * max(1, abs(loss_ref.item()))
    assert out[1].item() == (t != 255).sum().item()
    assert torch.equal(mask.cpu(), torch.max(logits.cpu(), 1)[1])                # bit-exact argmax (UCA:220)
    dx = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    dw, db = torch.empty(nc, C, 1, 1, device="cuda"), torch.empty(nc, device="cuda")
    call("unetca_outc_bwd", dt, ptr(g), ptr(gs), ptr(xd), C, ptr(dx), C, C, ptr(wd), nc, B, H * W, ptr(parts), ptr(dw),
         ptr(db), stream())
    assert relerr(dw.cpu(), w.grad) < 1e-4 and relerr(db.cpu(), b.grad) < 1e-4
    assert relerr(from_nhwc(dx), xr.grad) < TOL[dt]
    # all-ignored -> NaN, like torch
    td.fill_(255)
    call("unetca_cross_entropy", ptr(logits), ptr(td), nc, B, H * W, 255, None, None, None, ptr(parts), ptr(out), ptr(gs),
         stream())
    assert torch.isnan(out[0]).item()


def _conv_case(dt, impl, B, C, O, H, W, seed=4):
    rs = np.random.RandomState(seed)
    call("unetca_set_conv_impl", impl)
    x = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((O, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32))
    dy = torch.from_numpy(rs.standard_normal((B, O, H, W)).astype(np.float32))
    xr, wr, dyr = rounded(x, dt).requires_grad_(True), rounded(w, dt).requires_grad_(True), rounded(dy, dt)
    ref = F.conv2d(xr, wr, None, padding=1)
    ref.backward(dyr)
    xd, dyd, wdev = to_nhwc(x, dt), to_nhwc(dy, dt), w.cuda()
    wf = torch.empty(O, 9 * C, dtype=TDT[dt], device="cuda")
    wdg = torch.empty(C, 9 * O, dtype=TDT[dt], device="cuda")
    call("unetca_pack_conv3x3_weight", dt, ptr(wdev), ptr(wf), 9 * C, ptr(wdg), O, C, stream())
    y = torch.empty(B, H, W, O, dtype=TDT[dt], device="cuda")
    parts = parts_buf(B)
    n = cint()
    call("unetca_conv3x3_fwd", dt, ptr(xd), C, ptr(wf), 9 * C, ptr(y), O, B, H, W, C, O, ptr(parts), ctypes.byref(n), stream())
    tol = TOL[dt]
    got = from_nhwc(y)
    assert relerr(got, ref.detach()) < tol, "fwd"
    st = parts[: n.value * 2 * O].view(n.value, 2, O).sum(0).cpu()
    assert relerr(st[0], got.sum((0, 2, 3))) < 1e-3 and relerr(st[1], (got * got).sum((0, 2, 3))) < 1e-3, "stats"
    dx = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    call("unetca_conv3x3_fwd", dt, ptr(dyd), O, ptr(wdg), 9 * O, ptr(dx), C, B, H, W, O, C, None, None, stream())
    assert relerr(from_nhwc(dx), xr.grad) < tol, "dgrad"
    ws = torch.empty(8 * 1024 * 1024, device="cuda")
    dw = torch.empty(O, C, 3, 3, device="cuda")
    call("unetca_conv3x3_wgrad", dt, ptr(dyd), O, ptr(xd), C, ptr(ws), ws.numel(), B, H, W, C, O, ptr(dw), stream())
    assert relerr(dw.cpu(), wr.grad) < 2e-3, "wgrad"
    if impl == 0 and dt == BF16 and H % 2 == 0:
        # row-pair layout of the tcgen05 path (pixels on the MMA's N side, pair-packed filter): forward + stats + dgrad
        wfp = torch.empty(2 * O, 12 * C, dtype=TDT[dt], device="cuda")
        call("unetca_pack_conv3x3_pair", dt, ptr(wf), 9 * C, ptr(wfp), O, C, stream())
        y2 = torch.full((B, H, W, O), float("nan"), dtype=TDT[dt], device="cuda")
        call("unetca_conv3x3_fwd_paired", dt, ptr(xd), C, ptr(wfp), ptr(y2), O, B, H, W, C, O, ptr(parts), ctypes.byref(n),
             stream())
        got2 = from_nhwc(y2)
        assert relerr(got2, ref.detach()) < tol, "fwd (paired)"
        st = parts[: n.value * 2 * O].view(n.value, 2, O).sum(0).cpu()
        assert relerr(st[0], got2.sum((0, 2, 3))) < 1e-3 and relerr(st[1], (got2 * got2).sum((0, 2, 3))) < 1e-3, "stats (paired)"
        wdp = torch.empty(2 * C, 12 * O, dtype=TDT[dt], device="cuda")
        call("unetca_pack_conv3x3_pair", dt, ptr(wdg), 9 * O, ptr(wdp), C, O, stream())
        dx2 = torch.full((B, H, W, C), float("nan"), dtype=TDT[dt], device="cuda")
        call("unetca_conv3x3_fwd_paired", dt, ptr(dyd), O, ptr(wdp), ptr(dx2), C, B, H, W, O, C, None, None, stream())
        assert relerr(from_nhwc(dx2), xr.grad) < tol, "dgrad (paired)"
    if impl == 0 and dt == BF16 and O == 64 and C in (64, 128):
        # kw-stacked layout (N = 3 kw taps x 64 channels, shift-add epilogue, resident filter): forward + stats
        wkw = torch.empty(9 * C, 64, dtype=TDT[dt], device="cuda")
        call("unetca_pack_conv3x3_kw", dt, ptr(wf), 9 * C, ptr(wkw), C, stream())
        y3 = torch.full((B, H, W, O), float("nan"), dtype=TDT[dt], device="cuda")
        call("unetca_conv3x3_fwd_kw", dt, ptr(xd), C, ptr(wkw), ptr(y3), O, B, H, W, C, ptr(parts), ctypes.byref(n), stream())
        got3 = from_nhwc(y3)
        assert relerr(got3, ref.detach()) < tol, "fwd (kw)"
        st = parts[: n.value * 2 * O].view(n.value, 2, O).sum(0).cpu()
        assert relerr(st[0], got3.sum((0, 2, 3))) < 1e-3 and relerr(st[1], (got3 * got3).sum((0, 2, 3))) < 1e-3, "stats (kw)"
    if impl == 0 and dt == BF16 and C == 64 and O in (64, 128):
        # ... and as the dgrad of a layer with 64 input channels
        wkd = torch.empty(9 * O, 64, dtype=TDT[dt], device="cuda")
        call("unetca_pack_conv3x3_kw", dt, ptr(wdg), 9 * O, ptr(wkd), O, stream())
        dx3 = torch.full((B, H, W, C), float("nan"), dtype=TDT[dt], device="cuda")
        call("unetca_conv3x3_fwd_kw", dt, ptr(dyd), O, ptr(wkd), ptr(dx3), C, B, H, W, O, None, None, stream())
        assert relerr(from_nhwc(dx3), xr.grad) < tol, "dgrad (kw)"


@pytest.mark.parametrize("B,C,O,H,W", [(2, 64, 64, 16, 16), (1, 128, 64, 8, 24), (2, 64, 192, 4, 4)])
@pytest.mark.parametrize("dt", DTS)
def test_conv3x3_ffma(dt, B, C, O, H, W):
    _conv_case(dt, 1, B, C, O, H, W)


@pytest.mark.parametrize("B,C,O,H,W", [(2, 64, 64, 16, 16), (1, 128, 64, 8, 24), (2, 64, 192, 4, 4),
                                       (2, 128, 256, 32, 32), (3, 256, 128, 16, 48), (1, 64, 64, 128, 128),
                                       (2, 128, 128, 40, 24), (1, 192, 64, 48, 16), (2, 64, 128, 64, 64), (1, 64, 64, 34, 20),
                                       (1, 128, 384, 16, 16), (2, 64, 128, 36, 70), (1, 128, 64, 61, 33)])
def test_conv3x3_tcgen05(B, C, O, H, W):
    _conv_case(BF16, 0, B, C, O, H, W)


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (1, 64, 8), (3, 40, 72), (1, 130, 34), (2, 128, 160), (70, 16, 16)])
def test_conv3x3_rp64_resident_filter(B, H, W):
    """64 -> 64 channels through tc_conv3x3_rp64_kernel (row pairs, resident filter, haloed lattice tiles, direct global
    stores): forward + BatchNorm partial sums, and as a dgrad, vs conv2d on the bf16-rounded operands; partial tiles in both
    directions (W % 8, (H/2) % 32), untouched neighbours of a channel-slice output."""
    dt, C, O = BF16, 64, 64
    rs = np.random.RandomState(B * 1000 + H + W)
    call("unetca_set_conv_impl", 0)
    x = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((O, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32))
    dy = torch.from_numpy(rs.standard_normal((B, O, H, W)).astype(np.float32))
    xr, wr, dyr = rounded(x, dt).requires_grad_(True), rounded(w, dt), rounded(dy, dt)
    ref = F.conv2d(xr, wr, None, padding=1)
    ref.backward(dyr)
    xd, dyd = to_nhwc(x, dt), to_nhwc(dy, dt)
    wf = torch.empty(O, 9 * C, dtype=TDT[dt], device="cuda")
    wdg = torch.empty(C, 9 * O, dtype=TDT[dt], device="cuda")
    call("unetca_pack_conv3x3_weight", dt, ptr(w.cuda()), ptr(wf), 9 * C, ptr(wdg), O, C, stream())
    # output = the lower channel half of a wider buffer (ld = 128): the upper half must stay untouched
    buf = torch.full((B, H, W, 2 * O), 7.0, dtype=TDT[dt], device="cuda")
    y = buf[..., :O]
    parts = parts_buf(B)
    n = cint()
    call("unetca_conv3x3_fwd_rp64", dt, ptr(xd), C, ptr(wf), 9 * C, ptr(y), 2 * O, B, H, W, ptr(parts), ctypes.byref(n), stream())
    got = from_nhwc(y.contiguous())
    assert relerr(got, ref.detach()) < TOL[dt]
    assert bool((buf[..., O:] == 7.0).all())
    st = parts[: n.value * 2 * O].view(n.value, 2, O).sum(0).cpu()
    assert relerr(st[0], got.sum((0, 2, 3))) < 1e-3 and relerr(st[1], (got * got).sum((0, 2, 3))) < 1e-3
    dx = torch.full((B, H, W, C), float("nan"), dtype=TDT[dt], device="cuda")
    call("unetca_conv3x3_fwd_rp64", dt, ptr(dyd), O, ptr(wdg), 9 * O, ptr(dx), C, B, H, W, None, None, stream())
    assert relerr(from_nhwc(dx), xr.grad) < TOL[dt]
    # bit-identical to the per-tap row-pair kernel it replaces? (same products, fp32 accumulation order differs) -> close
    wp = torch.empty(2 * O, 12 * C, dtype=TDT[dt], device="cuda")
    call("unetca_pack_conv3x3_pair", dt, ptr(wf), 9 * C, ptr(wp), O, C, stream())
    y2 = torch.empty(B, H, W, O, dtype=TDT[dt], device="cuda")
    call("unetca_conv3x3_fwd_paired", dt, ptr(xd), C, ptr(wp), ptr(y2), O, B, H, W, C, O, None, None, stream())
    assert relerr(from_nhwc(y2), got) < 1e-2


@pytest.mark.parametrize("B,C,O,H,W", [(2, 64, 64, 16, 16), (3, 64, 64, 40, 72), (1, 64, 64, 130, 34), (2, 128, 128, 40, 24),
                                       (1, 256, 256, 33, 17), (3, 128, 128, 8, 8), (2, 512, 512, 16, 16)])
def test_conv3x3_dgrad_with_fused_bn_backward_statistics(B, C, O, H, W):
    """unetca_conv3x3_dgrad_bnstats: same dA1 as the plain dgrad, and its partial sums finalize to the same dgamma / dbeta /
    coefficients as the unetca_bn_bwd_reduce pass over (dA1, Y1) it replaces (autograd of UCA:82-84)."""
    dt = BF16
    rs = np.random.RandomState(C + H)
    call("unetca_set_conv_impl", 0)
    dy = torch.from_numpy(rs.standard_normal((B, O, H, W)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((O, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32))
    y1 = to_nhwc(torch.from_numpy((0.3 + rs.standard_normal((B, C, H, W))).astype(np.float32)), dt)
    dyd = to_nhwc(dy, dt)
    wf = torch.empty(O, 9 * C, dtype=TDT[dt], device="cuda")
    wdg = torch.empty(C, 9 * O, dtype=TDT[dt], device="cuda")
    call("unetca_pack_conv3x3_weight", dt, ptr(w.cuda()), ptr(wf), 9 * C, ptr(wdg), O, C, stream())
    scale = torch.from_numpy((0.5 + rs.rand(C)).astype(np.float32)).cuda()
    shift = torch.from_numpy((0.3 * rs.standard_normal(C)).astype(np.float32)).cuda()
    mean = torch.from_numpy((0.3 + 0.1 * rs.standard_normal(C)).astype(np.float32)).cuda()
    invstd = torch.from_numpy((0.8 + 0.4 * rs.rand(C)).astype(np.float32)).cuda()
    gamma = torch.from_numpy((1 + 0.1 * rs.standard_normal(C)).astype(np.float32)).cuda()
    # reference path: plain dgrad, then the reduce pass
    da_ref = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    call("unetca_conv3x3_fwd", dt, ptr(dyd), O, ptr(wdg), 9 * O, ptr(da_ref), C, B, H, W, O, C, None, None, stream())
    parts = parts_buf(B)
    n = cint()
    call("unetca_bn_bwd_reduce", dt, ptr(da_ref), C, ptr(y1), C, B, H * W, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd), None,
         None, ptr(parts), ctypes.byref(n), stream())
    out_ref = [torch.empty(C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(3, C, device="cuda")]
    call("unetca_bn_bwd_finalize", ptr(parts), n.value, C, B * H * W, ptr(gamma), ptr(invstd), ptr(out_ref[0]), ptr(out_ref[1]),
         ptr(out_ref[2]), stream())
    # fused
    da = torch.full((B, H, W, C), float("nan"), dtype=TDT[dt], device="cuda")
    parts2 = parts_buf(B)
    n2 = cint()
    call("unetca_conv3x3_dgrad_bnstats", dt, ptr(dyd), O, ptr(wdg), 9 * O, ptr(da), C, B, H, W, O, C, ptr(y1), C, ptr(scale), ptr(shift),
         ptr(mean), ptr(parts2), ctypes.byref(n2), stream())
    out = [torch.empty(C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(3, C, device="cuda")]
    call("unetca_bn_bwd_finalize", ptr(parts2), n2.value, C, B * H * W, ptr(gamma), ptr(invstd), ptr(out[0]), ptr(out[1]), ptr(out[2]),
         stream())
    if C == 64:
        assert relerr(da.float().cpu(), da_ref.float().cpu()) < 1e-2        # rp64 vs the kernel unetca_conv3x3_fwd picks for O = 64
    else:
        assert torch.equal(da, da_ref)                                      # same haloed kernel, same arithmetic
    for a, b in zip(out, out_ref):
        # sums of the same bf16 values in a different order (per-thread columns vs per-pixel-row blocks): fp32 round-off,
        # measured against the mass that was summed
        assert relerr(a.cpu(), b.cpu()) < (2e-2 if C == 64 else 1e-3)


@pytest.mark.parametrize("B,C,O,H,W,layout", [(3, 128, 128, 40, 72, 0), (2, 64, 256, 16, 48, 0), (5, 128, 128, 8, 8, 0),
                                              (3, 64, 64, 40, 72, 1), (2, 64, 64, 128, 160, 1), (70, 64, 64, 16, 16, 1),
                                              (1, 128, 64, 24, 40, 2), (3, 64, 64, 40, 72, 3), (70, 64, 64, 16, 16, 3)])
def test_conv3x3_bnrelu_squeeze_epilogue(B, C, O, H, W, layout):
    """Inference form of a block's convolution: eval-mode BatchNorm + ReLU in the epilogue (UCA:83-90 with running
    statistics) and, for a block's second convolution, the SE squeeze (UCA:65) as per-image partial channel sums."""
    dt = BF16
    rs = np.random.RandomState(17)
    call("unetca_set_conv_impl", 0)
    x = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((O, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32))
    scale = torch.from_numpy((0.5 + rs.rand(O)).astype(np.float32))
    shift = torch.from_numpy((0.3 * rs.standard_normal(O)).astype(np.float32))
    ref = F.relu(F.conv2d(rounded(x, dt), rounded(w, dt), None, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    xd, wdev = to_nhwc(x, dt), w.cuda()
    wf = torch.empty(O, 9 * C, dtype=TDT[dt], device="cuda")
    wdg = torch.empty(C, 9 * O, dtype=TDT[dt], device="cuda")
    call("unetca_pack_conv3x3_weight", dt, ptr(wdev), ptr(wf), 9 * C, ptr(wdg), O, C, stream())
    wt = wf
    if layout == 1:
        wt = torch.empty(2 * O, 12 * C, dtype=TDT[dt], device="cuda")
        call("unetca_pack_conv3x3_pair", dt, ptr(wf), 9 * C, ptr(wt), O, C, stream())
    elif layout == 2:
        wt = torch.empty(9 * C, 64, dtype=TDT[dt], device="cuda")
        call("unetca_pack_conv3x3_kw", dt, ptr(wf), 9 * C, ptr(wt), C, stream())
    y = torch.full((B, H, W, O), float("nan"), dtype=TDT[dt], device="cuda")
    sq = None
    if layout != 2:
        sq = torch.zeros(B * unetca_b200._lib.load().unetca_num_sms() * O, device="cuda")
    n = cint()
    scale_d, shift_d = scale.cuda(), shift.cuda()
    call("unetca_conv3x3_bnrelu_fwd", dt, ptr(xd), C, ptr(wt), layout, ptr(y), O, B, H, W, C, O, ptr(scale_d),
         ptr(shift_d), ptr(sq), ctypes.byref(n), stream())
    got = from_nhwc(y)
    assert relerr(got, ref) < TOL[dt]
    if sq is not None:
        assert 0 < n.value <= unetca_b200._lib.load().unetca_num_sms()
        sums = sq[: B * n.value * O].view(B, n.value, O).sum(1).cpu()
        want = got.sum((2, 3))                                   # of the stored (bf16) activations
        assert relerr(sums, want) < 1e-4
        # ... and they feed unetca_se_fc unchanged: its pooled output is the per-image channel mean
        Cr = O // 16
        w1 = torch.from_numpy((rs.standard_normal((Cr, O)) / np.sqrt(O)).astype(np.float32)).cuda()
        w2 = torch.from_numpy((rs.standard_normal((O, Cr)) / np.sqrt(Cr)).astype(np.float32)).cuda()
        pl, z, sg = torch.empty(B, O, device="cuda"), torch.empty(B, Cr, device="cuda"), torch.empty(B, O, device="cuda")
        call("unetca_se_fc", ptr(sq), n.value, B, O, Cr, H * W, ptr(w1), ptr(w2), ptr(pl), ptr(z), ptr(sg), stream())
        mean = want / (H * W)
        assert relerr(pl.cpu(), mean) < 1e-4
        assert relerr(sg.cpu(), torch.sigmoid(F.relu(mean @ w1.cpu().t()) @ w2.cpu().t())) < 1e-4


def _convT_case(dt, impl, B, Cin, Cout, h, w_, seed=5):
    rs = np.random.RandomState(seed)
    call("unetca_set_conv_impl", impl)
    x = torch.from_numpy(rs.standard_normal((B, Cin, h, w_)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((Cin, Cout, 2, 2)) / np.sqrt(Cin)).astype(np.float32))
    b = torch.from_numpy((0.1 * rs.standard_normal(Cout)).astype(np.float32))
    dy = torch.from_numpy(rs.standard_normal((B, Cout, 2 * h, 2 * w_)).astype(np.float32))
    xr, wr, dyr = rounded(x, dt).requires_grad_(True), rounded(w, dt).requires_grad_(True), rounded(dy, dt)
    ref = F.conv_transpose2d(xr, wr, b, stride=2)
    ref.backward(dyr)
    xd, wdev, bd = to_nhwc(x, dt), w.cuda(), b.cuda()
    wf = torch.empty(4 * Cout, Cin, dtype=TDT[dt], device="cuda")
    wdg = torch.empty(Cin, 4 * Cout, dtype=TDT[dt], device="cuda")
    call("unetca_pack_convT_weight", dt, ptr(wdev), ptr(wf), ptr(wdg), Cin, Cout, stream())
    cat = torch.zeros(B, 2 * h, 2 * w_, 2 * Cout, dtype=TDT[dt], device="cuda")      # write the upper channel half
    call("unetca_convT2x2_fwd", dt, ptr(xd), Cin, ptr(wf), ptr(bd), ptr(cat[..., Cout:]), 2 * Cout, B, h, w_, Cin, Cout,
         stream())
    tol = TOL[dt]
    assert relerr(from_nhwc(cat[..., Cout:]), ref.detach()) < tol, "fwd"
    assert cat[..., :Cout].abs().max().item() == 0
    dcat = torch.zeros(B, 2 * h, 2 * w_, 2 * Cout, dtype=TDT[dt], device="cuda")
    dcat[..., Cout:] = to_nhwc(dy, dt)
    dx = torch.empty(B, h, w_, Cin, dtype=TDT[dt], device="cuda")
    call("unetca_convT2x2_dgrad", dt, ptr(dcat[..., Cout:]), 2 * Cout, ptr(wdg), ptr(dx), Cin, B, h, w_, Cin, Cout, stream())
    assert relerr(from_nhwc(dx), xr.grad) < tol, "dgrad"
    ws = torch.empty(8 * 1024 * 1024, device="cuda")
    dw = torch.empty(Cin, Cout, 2, 2, device="cuda")
    call("unetca_convT2x2_wgrad", dt, ptr(xd), Cin, ptr(dcat[..., Cout:]), 2 * Cout, ptr(ws), ws.numel(), B, h, w_, Cin, Cout,
         ptr(dw), stream())
    assert relerr(dw.cpu(), wr.grad) < 2e-3, "wgrad"
    dbias = torch.empty(Cout, device="cuda")
    call("unetca_chan_sum", dt, ptr(dcat[..., Cout:]), 2 * Cout, Cout, B * 4 * h * w_, ptr(parts_buf(B)), ptr(dbias), stream())
    assert relerr(dbias.cpu(), dyr.sum((0, 2, 3))) < 1e-3, "dbias"


@pytest.mark.parametrize("dt", DTS)
def test_convT_ffma(dt):
    _convT_case(dt, 1, 2, 128, 64, 8, 8)


@pytest.mark.parametrize("B,Cin,Cout,h,w_", [(2, 128, 64, 8, 8), (1, 256, 128, 16, 32), (3, 1024, 512, 2, 2),
                                             # the dedicated pixels-on-N forward kernel (h >= 16, w >= 8): partial tiles in H / W / both
                                             (2, 128, 64, 40, 24), (1, 256, 128, 33, 17), (2, 512, 256, 64, 8), (1, 1024, 512, 32, 32)])
def test_convT_tcgen05(B, Cin, Cout, h, w_):
    _convT_case(BF16, 0, B, Cin, Cout, h, w_)


@pytest.mark.parametrize("Cin,h,w_", [(128, 64, 64), (256, 33, 17), (512, 32, 16)])
def test_convT_forward_kernels_agree(Cin, h, w_):
    """The dedicated ConvTranspose forward kernel (sub-pixel channels on M, direct stores) against the generic pixels-on-M one, into
    a dense tensor and into the channel half of a concat buffer: same K order, same rounding -> identical bits."""
    dt, B, Cout = BF16, 2, Cin // 2
    rs = np.random.RandomState(Cin + h)
    call("unetca_set_conv_impl", 0)
    x = to_nhwc(torch.from_numpy(rs.standard_normal((B, Cin, h, w_)).astype(np.float32)), dt)
    wf = torch.from_numpy((rs.standard_normal((4 * Cout, Cin)) / np.sqrt(Cin)).astype(np.float32)).cuda().to(TDT[dt])
    bias = torch.from_numpy(rs.standard_normal(Cout).astype(np.float32)).cuda()
    res = []
    try:
        for pix in (0, 1):
            unetca_b200._lib.load().unetca_tc_set_convT_pix(pix)
            dense = torch.full((B, 2 * h, 2 * w_, Cout), float("nan"), dtype=TDT[dt], device="cuda")
            cat = torch.zeros(B, 2 * h, 2 * w_, 2 * Cout, dtype=TDT[dt], device="cuda")
            call("unetca_convT2x2_fwd", dt, ptr(x), Cin, ptr(wf), ptr(bias), ptr(dense), Cout, B, h, w_, Cin, Cout, stream())
            call("unetca_convT2x2_fwd", dt, ptr(x), Cin, ptr(wf), ptr(bias), ptr(cat[..., Cout:]), 2 * Cout, B, h, w_, Cin, Cout, stream())
            assert cat[..., :Cout].abs().max().item() == 0 and torch.equal(cat[..., Cout:], dense)
            res.append(dense)
    finally:
        unetca_b200._lib.load().unetca_tc_set_convT_pix(1)
    assert torch.equal(res[0], res[1])


@pytest.mark.parametrize("dt,impl", [(F32, 1), (BF16, 1), (BF16, 0)])
def test_first_conv_im2col(dt, impl):
    """K = 9*Cin = 27 first conv: im2col rows + NT GEMM forward, TN GEMM weight gradient."""
    rs = np.random.RandomState(6)
    call("unetca_set_conv_impl", impl)
    B, Cin, O, H, W = 2, 3, 64, 16, 24
    x = torch.from_numpy(rs.standard_normal((B, Cin, H, W)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((O, Cin, 3, 3)) / np.sqrt(27)).astype(np.float32))
    dy = torch.from_numpy(rs.standard_normal((B, O, H, W)).astype(np.float32))
    xr, wr, dyr = rounded(x, dt), rounded(w, dt).requires_grad_(True), rounded(dy, dt)
    ref = F.conv2d(xr, wr, None, padding=1)
    ref.backward(dyr)
    Kpad = 64
    col = torch.empty(B * H * W, Kpad, dtype=TDT[dt], device="cuda")
    call("unetca_im2col3x3_nchw", dt, ptr(x.cuda()), ptr(col), B, Cin, H, W, Kpad, stream())
    wf = torch.empty(O, Kpad, dtype=TDT[dt], device="cuda")
    call("unetca_pack_conv3x3_weight", dt, ptr(w.cuda()), ptr(wf), Kpad, None, O, Cin, stream())
    y = torch.empty(B, H, W, O, dtype=TDT[dt], device="cuda")
    parts = parts_buf(B)
    n = cint()
    call("unetca_gemm_nt", dt, ptr(col), Kpad, ptr(wf), Kpad, ptr(y), O, B * H * W, O, Kpad, ptr(parts), ctypes.byref(n),
         stream())
    assert relerr(from_nhwc(y), ref.detach()) < TOL[dt]
    dyd = to_nhwc(dy, dt)
    ws = torch.empty(4 * 1024 * 1024, device="cuda")
    dw = torch.empty(O, Cin, 3, 3, device="cuda")
    call("unetca_im2col_wgrad", dt, ptr(dyd), O, ptr(col), Kpad, ptr(ws), ws.numel(), B * H * W, Cin, O, ptr(dw), stream())
    assert relerr(dw.cpu(), wr.grad) < 2e-3


@pytest.mark.parametrize("B,Cin,H,W", [(2, 3, 16, 24), (1, 1, 34, 20), (3, 3, 64, 64)])
def test_first_conv_pixel_pairs(B, Cin, H, W):
    """First conv through the row-pair layout: pixel-pair im2col + pair-packed filter + one pixels-on-N GEMM (forward,
    BatchNorm partial sums) and the weight gradient from the same pair rows, vs conv2d on the bf16-rounded operands."""
    rs = np.random.RandomState(16)
    call("unetca_set_conv_impl", 0)
    dt, O = BF16, 64
    x = torch.from_numpy(rs.standard_normal((B, Cin, H, W)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((O, Cin, 3, 3)) / np.sqrt(9 * Cin)).astype(np.float32))
    dy = torch.from_numpy(rs.standard_normal((B, O, H, W)).astype(np.float32))
    xr, wr, dyr = rounded(x, dt), rounded(w, dt).requires_grad_(True), rounded(dy, dt)
    ref = F.conv2d(xr, wr, None, padding=1)
    ref.backward(dyr)
    colp = torch.empty(B * (H // 2) * W, 64, dtype=TDT[dt], device="cuda")
    call("unetca_im2col_pairs", dt, ptr(x.cuda()), ptr(colp), B, Cin, H, W, stream())
    wp = torch.empty(2 * O, 64, dtype=TDT[dt], device="cuda")
    call("unetca_pack_first_pairs", dt, ptr(w.cuda()), ptr(wp), O, Cin, stream())
    y = torch.full((B, H, W, O), float("nan"), dtype=TDT[dt], device="cuda")
    parts = parts_buf(B)
    n = cint()
    call("unetca_first_pairs_fwd", dt, ptr(colp), ptr(wp), ptr(y), O, B, H, W, O, ptr(parts), ctypes.byref(n), stream())
    got = from_nhwc(y)
    assert relerr(got, ref.detach()) < TOL[dt]
    st = parts[: n.value * 2 * O].view(n.value, 2, O).sum(0).cpu()
    assert relerr(st[0], got.sum((0, 2, 3))) < 1e-3 and relerr(st[1], (got * got).sum((0, 2, 3))) < 1e-3
    dyd = to_nhwc(dy, dt)
    ws = torch.empty(4 * 1024 * 1024, device="cuda")
    dw = torch.empty(O, Cin, 3, 3, device="cuda")
    call("unetca_first_pairs_wgrad", dt, ptr(dyd), O, ptr(colp), ptr(ws), ws.numel(), B, H, W, Cin, O, ptr(dw), stream())
    assert relerr(dw.cpu(), wr.grad) < 2e-3


def test_metrics_on_device_match_reference(golden_dir):
    """f-1: compute_metrics (UCA:214-269) from the on-device label-by-prediction table == the reference's numbers."""
    import os
    from oracle import metrics_port
    from unetca_b200 import metrics
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    for name, logits, masks, nc in metrics_port.metric_cases():
        counts = metrics.confusion_counts(logits.cuda(), masks.cuda(), nc).cpu().numpy()
        preds = torch.max(logits, 1)[1]
        valid = masks != 255
        for r in range(nc + 1):
            for c in range(nc):
                lab = (masks == r) if r < nc else ((masks < 0) | (masks >= nc))
                assert counts[r, c] == int((lab & valid & (preds == c)).sum()), (name, r, c)      # exact integers
        m = metrics.compute_metrics(logits.cuda(), masks.cuda(), nc)
        np.testing.assert_allclose([m["acc"], m["miou"], m["mpa"], m["mf1"]], g[name], rtol=0, atol=1e-15, err_msg=name)
    # a full-size batch (configs[1] shape): counts add up and agree with torch on the device
    lo = torch.randn(8, 2, 512, 512, device="cuda")
    ma = torch.randint(0, 2, (8, 512, 512), device="cuda")
    ma[torch.rand(8, 512, 512, device="cuda") < 0.01] = 255
    counts = metrics.confusion_counts(lo, ma, 2)
    pr = torch.max(lo, 1)[1]
    for r in range(2):
        for c in range(2):
            assert counts[r, c].item() == ((ma == r) & (pr == c)).sum().item()
    assert counts.sum().item() == (ma != 255).sum().item()


def test_device_input_pipeline_matches_reference(golden_dir):
    """f-3: uint8 batches -> pinned H2D on a side stream -> unetca_prep_u8 == the reference's ToTensor/Normalize and
    mask ToTensor().long() tensors, bit for bit (tests/golden/voc_mini_expected.npz from the reference's dataset)."""
    import os
    from torch.utils.data import DataLoader
    from unetca_b200 import data
    root = os.path.join(golden_dir, "voc_mini")
    g = np.load(os.path.join(golden_dir, "voc_mini_expected.npz"))
    ds = data.VOCSegDataset(root, 32, image_set="train", raw=True)
    loader = DataLoader(ds, batch_size=2, shuffle=False, num_workers=0)
    seen = 0
    for bi, (x, y) in enumerate(data.DevicePrefetcher(loader, torch.device("cuda"))):
        assert x.is_cuda and x.dtype == torch.float32 and x.shape == (2, 1, 32, 32)
        assert y.is_cuda and y.dtype == torch.int64 and y.shape == (2, 32, 32)
        for k in range(2):
            name = ds.ids[bi * 2 + k]
            assert np.array_equal(x[k].cpu().numpy(), g[f"img:{name}"]), name
            assert np.array_equal(y[k].cpu().numpy(), g[f"mask:{name}"]), name
            seen += 1
    assert seen == 4
    # odd sizes / unaligned tails and every byte value
    u = torch.arange(256, dtype=torch.uint8).repeat(5)[:1237].reshape(1, 1, 1237).cuda()
    xi, yi = data.prep_u8(u, u)
    ref = ((u.cpu().float() / 255.0) - 0.5) / 0.5
    assert torch.equal(xi.cpu()[:, 0], ref) and torch.equal(yi.cpu(), (u.cpu().float() / 255.0).long())


def _guarded(shape, dtype, pad=4096):
    """A tensor view of `shape` carved out of the middle of a sentinel-filled buffer, plus a checker that the guard
    bands on both sides are untouched (compute-sanitizer is not available on the GPU pool)."""
    n = int(np.prod(shape))
    buf = torch.full((n + 2 * pad,), 123.0 if dtype.is_floating_point else 77, dtype=dtype, device="cuda")
    view = buf[pad:pad + n].view(*shape)

    def intact():
        ref = 123.0 if dtype.is_floating_point else 77
        return bool((buf[:pad] == ref).all().item() and (buf[pad + n:] == ref).all().item())
    return view, intact


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 7, 13), (1, 192, 5, 6), (3, 128, 9, 4)])
def test_no_out_of_bounds_writes_on_odd_shapes(B, C, H, W):
    """Guard-band check of the HBM-bound kernels on odd extents: nothing is written outside the output tensors."""
    dt = BF16
    rs = np.random.RandomState(3)
    tdt = TDT[dt]
    y = torch.from_numpy(rs.standard_normal((B, H, W, C)).astype(np.float32)).to(tdt).cuda()
    d = torch.from_numpy(rs.standard_normal((B, H, W, C)).astype(np.float32)).to(tdt).cuda()
    sc, sh = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
    mean, invstd = torch.randn(C, device="cuda") * 0.1, torch.rand(C, device="cuda") + 0.5
    s, dp = torch.rand(B, C, device="cuda"), torch.randn(B, C, device="cuda")
    coef = torch.rand(3, C, device="cuda")
    checks = []
    out, ok = _guarded((B, H, W, C), tdt); checks.append(ok)
    call("unetca_bn_relu", dt, ptr(y), C, ptr(out), C, B, H * W, C, ptr(sc), ptr(sh), None, None, stream())
    out2, ok = _guarded((B, H, W, C), tdt); checks.append(ok)
    call("unetca_se_scale_pool", dt, ptr(y), C, ptr(out2), C, None, 0, None, B, H, W, C, ptr(sc), ptr(sh), ptr(s), stream())
    pooled, ok = _guarded((B, H // 2, W // 2, C), tdt); checks.append(ok)
    pos, ok = _guarded((B, H // 2, W // 2, C), torch.uint8); checks.append(ok)
    call("unetca_maxpool2x2", dt, ptr(y), C, ptr(pooled), C, ptr(pos), None, B, H, W, C, stream())
    dx, ok = _guarded((B, H, W, C), tdt); checks.append(ok)
    call("unetca_pool_bwd_add", dt, ptr(d), C, ptr(pooled), C, ptr(pos), ptr(dx), C, B, H, W, C, stream())
    dy, ok = _guarded((B, H, W, C), tdt); checks.append(ok)
    call("unetca_bn_bwd_apply", dt, ptr(d), C, ptr(y), C, ptr(dy), C, B, H * W, C, ptr(sc), ptr(sh), ptr(mean), ptr(invstd),
         ptr(s), ptr(dp), ptr(coef), stream())
    up, ok = _guarded((B, H + 1, W + 1, C), tdt); checks.append(ok)
    call("unetca_resize_bilinear_fwd", dt, ptr(y), C, H, W, ptr(up), C, H + 1, W + 1, B, C, stream())
    dn, ok = _guarded((B, H, W, C), tdt); checks.append(ok)
    call("unetca_resize_bilinear_bwd", dt, ptr(up), C, H + 1, W + 1, ptr(dn), C, H, W, B, C, stream())
    parts, ok = _guarded((_lib_max_parts(B) * 2 * C,), torch.float32); checks.append(ok)
    n = cint()
    call("unetca_se_bn_bwd_reduce", dt, ptr(d), C, ptr(y), C, B, H * W, C, ptr(sc), ptr(sh), ptr(mean), ptr(parts),
         ctypes.byref(n), stream())
    assert n.value * B <= _lib_max_parts(B)
    torch.cuda.synchronize()
    assert all(ok() for ok in checks)


def _lib_max_parts(B):
    from unetca_b200 import _lib
    return _lib.load().unetca_max_parts(B)


@pytest.mark.parametrize("dt", DTS)
def test_encoder_backward_with_on_the_fly_unpool(dt):
    """se_bn_bwd_reduce_pool / bn_bwd_apply_pool (dO rebuilt per quad from skip gradient + max-pool routing) against the
    two-step path pool_bwd_add -> se_bn_bwd_reduce / bn_bwd_apply; the fused kernels keep dO in fp32, the two-step path
    rounds it to the storage type in between, hence the tolerance in bf16."""
    rs = np.random.RandomState(12)
    B, C, H, W = 2, 128, 12, 16
    tdt = TDT[dt]
    y = torch.from_numpy(rs.standard_normal((B, H, W, C)).astype(np.float32)).to(tdt).cuda()
    sg = torch.from_numpy(rs.standard_normal((B, H, W, 2 * C)).astype(np.float32)).to(tdt).cuda()[..., :C]   # strided view
    dpl = torch.from_numpy(rs.standard_normal((B, H // 2, W // 2, C)).astype(np.float32)).to(tdt).cuda()
    pos = torch.from_numpy(rs.randint(0, 4, (B, H // 2, W // 2, C)).astype(np.uint8)).cuda()
    sc, sh = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.2
    mean, invstd = torch.randn(C, device="cuda") * 0.2, torch.rand(C, device="cuda") + 0.5
    s, dp = torch.rand(B, C, device="cuda"), torch.randn(B, C, device="cuda")
    coef = torch.rand(3, C, device="cuda")
    dcur = torch.empty(B, H, W, C, dtype=tdt, device="cuda")
    call("unetca_pool_bwd_add", dt, ptr(sg), 2 * C, ptr(dpl), C, ptr(pos), ptr(dcur), C, B, H, W, C, stream())
    pa, pb = parts_buf(B, 4096), parts_buf(B, 4096)
    na, nb = cint(), cint()
    call("unetca_se_bn_bwd_reduce", dt, ptr(dcur), C, ptr(y), C, B, H * W, C, ptr(sc), ptr(sh), ptr(mean), ptr(pa),
         ctypes.byref(na), stream())
    call("unetca_se_bn_bwd_reduce_pool", dt, ptr(sg), 2 * C, ptr(dpl), C, ptr(pos), ptr(y), C, B, H, W, C, ptr(sc), ptr(sh),
         ptr(mean), ptr(pb), ctypes.byref(nb), stream())
    ra = pa[: B * na.value * 2 * C].view(B, na.value, 2, C).sum(1)
    rb = pb[: B * nb.value * 2 * C].view(B, nb.value, 2, C).sum(1)
    assert relerr(rb, ra) < (1e-5 if dt == F32 else 1e-2)
    dya = torch.empty(B, H, W, C, dtype=tdt, device="cuda")
    dyb = torch.empty(B, H, W, C, dtype=tdt, device="cuda")
    call("unetca_bn_bwd_apply", dt, ptr(dcur), C, ptr(y), C, ptr(dya), C, B, H * W, C, ptr(sc), ptr(sh), ptr(mean), ptr(invstd),
         ptr(s), ptr(dp), ptr(coef), stream())
    call("unetca_bn_bwd_apply_pool", dt, ptr(sg), 2 * C, ptr(dpl), C, ptr(pos), ptr(y), C, ptr(dyb), C, B, H, W, C, ptr(sc),
         ptr(sh), ptr(mean), ptr(invstd), ptr(s), ptr(dp), ptr(coef), stream())
    assert relerr(dyb.float(), dya.float()) < (1e-5 if dt == F32 else 1.5e-2)


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adam_multi_tensor_matches_torch(wd):
    """optimizer.step() (UCA:346) with optim.Adam (UCA:466): 6 steps on tensors of awkward sizes against torch.optim.Adam
    on the CPU; fp32 both sides, differences are rounding only (fused multiply-adds): 2e-6 of the largest value."""
    from unetca_b200 import optim as uoptim
    rs = np.random.RandomState(3)
    shapes = [(1,), (7,), (64,), (2, 64, 1, 1), (4096,), (4097,), (3, 4096), (64, 3, 3, 3), (33, 5, 7), (100000,)] + [(64,)] * 70
    ref = [torch.nn.Parameter(torch.from_numpy(rs.standard_normal(sh).astype(np.float32))) for sh in shapes]
    own = [torch.nn.Parameter(r.detach().clone().cuda()) for r in ref]
    ropt = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    opt = uoptim.Adam(own, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    for step in range(6):
        for r, o in zip(ref, own):
            g = torch.from_numpy(rs.standard_normal(tuple(r.shape)).astype(np.float32)) * (10.0 ** (step - 3))
            r.grad = g
            o.grad = g.cuda() if not (step == 2 and r.numel() == 7) else None      # one skipped tensor in one step
            if o.grad is None:
                r.grad = None
        ropt.step()
        opt.step()
    for r, o in zip(ref, own):
        if r.numel() == 7:
            continue          # skipped once: torch keeps a per-tensor step count, this optimizer one per group (documented)
        assert relerr(o.detach().cpu(), r.detach()) < 2e-6
        assert relerr(opt.state[o]["exp_avg"].cpu(), ropt.state[r]["exp_avg"]) < 2e-6
        assert relerr(opt.state[o]["exp_avg_sq"].cpu(), ropt.state[r]["exp_avg_sq"]) < 2e-6
    # state_dict round trip into torch.optim.Adam and back
    sd = opt.state_dict()
    assert float(sd["state"][0]["step"]) == 6.0
    topt = torch.optim.Adam([torch.nn.Parameter(o.detach().clone()) for o in own], lr=1e-2, betas=(0.9, 0.99), weight_decay=wd)
    import copy
    topt.load_state_dict(copy.deepcopy(sd))          # (torch's load_state_dict aliases same-device state tensors)
    opt2 = uoptim.Adam([torch.nn.Parameter(o.detach().clone()) for o in own], lr=1e-2, betas=(0.9, 0.99), weight_decay=wd)
    opt2.load_state_dict(topt.state_dict())
    for a, b, c in zip(own, topt.param_groups[0]["params"], opt2.param_groups[0]["params"]):
        if a.numel() == 7:
            continue
        g = torch.ones_like(a)
        a.grad, b.grad, c.grad = g, g.clone(), g.clone()
    opt.step(); topt.step(); opt2.step()
    for a, b, c in zip(own, topt.param_groups[0]["params"], opt2.param_groups[0]["params"]):
        if a.numel() == 7:
            continue
        assert torch.equal(a, c)
        assert relerr(a, b) < 2e-6
    with pytest.raises(RuntimeError):
        cpu_p = torch.nn.Parameter(torch.zeros(4)); cpu_p.grad = torch.ones(4)
        uoptim.Adam([cpu_p]).step()


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("O,C", [(64, 96), (128, 64), (32, 32)])
def test_adam_conv3x3_emits_packed_filters(dt, O, C):
    """The conv-filter Adam kernel: same step as the multi-tensor kernel, and the packed operand copies it writes are
    bit-identical to unetca_pack_conv3x3_weight of the stepped weights."""
    rs = np.random.RandomState(5)
    mk = lambda sc=1.0: torch.from_numpy((sc * rs.standard_normal((O, C, 3, 3))).astype(np.float32)).cuda()
    p, g, m, v = mk(), mk(0.1), mk(0.05), mk(0.02).abs()
    p2, m2, v2 = p.clone(), m.clone(), v.clone()
    hyper = torch.zeros(8, device="cuda")
    hyper[0] = 4
    call("unetca_adam_tick", ptr(hyper), 1e-3, 0.9, 0.999, 1e-8, 0.0, stream())
    assert hyper[0].item() == 5.0
    assert abs(hyper[1].item() - 1e-3 / (1 - 0.9 ** 5)) < 1e-9 and abs(hyper[2].item() - (1 - 0.999 ** 5) ** 0.5) < 1e-8
    wf = torch.full((O, 9 * C), float("nan"), dtype=TDT[dt], device="cuda")
    wd = torch.full((C, 9 * O), float("nan"), dtype=TDT[dt], device="cuda")
    tab = np.asarray([(ptr(p), ptr(g), ptr(m), ptr(v), ptr(wf), ptr(wd), O, C)], dtype=np.int64)
    assert call("unetca_adam_step_conv3x3", dt, tab.ctypes.data_as(ctypes.c_void_p), 1, ptr(hyper), stream()) == 1
    tab2 = np.asarray([(ptr(p2), ptr(g), ptr(m2), ptr(v2), p2.numel())], dtype=np.int64)
    assert call("unetca_adam_step", tab2.ctypes.data_as(ctypes.c_void_p), 1, ptr(hyper), stream()) == 1
    assert torch.equal(p, p2) and torch.equal(m, m2) and torch.equal(v, v2)
    wf_ref = torch.empty_like(wf)
    wd_ref = torch.empty_like(wd)
    call("unetca_pack_conv3x3_weight", dt, ptr(p), ptr(wf_ref), 9 * C, ptr(wd_ref), O, C, stream())
    assert torch.equal(wf.view(torch.int16 if dt == BF16 else torch.int32), wf_ref.view(torch.int16 if dt == BF16 else torch.int32))
    assert torch.equal(wd.view(torch.int16 if dt == BF16 else torch.int32), wd_ref.view(torch.int16 if dt == BF16 else torch.int32))


@pytest.mark.parametrize("B,C,HW,se", [(2, 64, 16 * 16, False), (3, 128, 37 * 45, True), (2, 1024, 2 * 2, True), (1, 64, 1, False),
                                       (2, 256, 640, True), (5, 512, 4 * 5, False)])
def test_bn_bwd_apply_stream_equals_register_kernel(B, C, HW, se):
    """The shared-memory-streamed form of the ReLU+BN(+SE) backward apply pass (cp.async.bulk ring) must give the very
    bits of the register kernel — whole tiles, ragged tails, ranges shorter than one tile."""
    rs = np.random.RandomState(9)
    mk = lambda *sh: torch.from_numpy(rs.standard_normal(sh).astype(np.float32)).cuda()
    y, d = mk(B, HW, C).bfloat16(), mk(B, HW, C).bfloat16()
    scale, shift, mean, invstd = mk(C), mk(C), mk(C), mk(C).abs() + 0.5
    coef = mk(3, C)
    s_, dp = (torch.sigmoid(mk(B, C)), mk(B, C)) if se else (None, None)
    outs = []
    try:
        for tiles in (0, 1, 3, 16):
            unetca_b200._lib.load().unetca_set_tuning(3, tiles)
            dy = torch.full((B, HW, C), float("nan"), dtype=torch.bfloat16, device="cuda")
            guard = torch.full((B * HW * C + 4096,), 7.0, dtype=torch.bfloat16, device="cuda")     # guard band behind a second output
            dy2 = guard[: B * HW * C].view(B, HW, C)
            for o in (dy, dy2):
                call("unetca_bn_bwd_apply", BF16, ptr(d), C, ptr(y), C, ptr(o), C, B, HW, C, ptr(scale), ptr(shift), ptr(mean),
                     ptr(invstd), ptr(s_), ptr(dp), ptr(coef), stream())
            assert torch.equal(dy.view(torch.int16), dy2.view(torch.int16))
            assert torch.all(guard[B * HW * C:] == 7.0)
            assert not torch.isnan(dy.float()).any()
            outs.append(dy)
    finally:
        unetca_b200._lib.load().unetca_set_tuning(3, DEFAULT_APPLY_STREAM)
    for o in outs[1:]:
        assert torch.equal(o.view(torch.int16), outs[0].view(torch.int16))


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 8, 128), (3, 64, 6, 70), (2, 128, 12, 16), (1, 256, 4, 34), (2, 256, 18, 16), (1, 64, 2, 2)])
def test_pool_backward_stream_equals_register_kernels(B, C, H, W):
    """The quad-row shared-memory stream of se_bn_bwd_reduce_pool / bn_bwd_apply_pool (TMA tensor copies, sg a channel
    slice of a wider buffer) against the register kernels: apply bit-identical, sums equal up to fp32 summation order."""
    rs = np.random.RandomState(21)
    mk = lambda *sh: torch.from_numpy(rs.standard_normal(sh).astype(np.float32)).cuda()
    y = mk(B, H, W, C).bfloat16()
    sg = mk(B, H, W, 2 * C).bfloat16()[..., C:]                       # strided view (the skip half of a concat gradient)
    dpl = mk(B, H // 2, W // 2, C).bfloat16()
    pos = torch.from_numpy(rs.randint(0, 4, (B, H // 2, W // 2, C)).astype(np.uint8)).cuda()
    sc, sh = torch.rand(C, device="cuda") + 0.5, mk(C) * 0.2
    mean, invstd = mk(C) * 0.2, torch.rand(C, device="cuda") + 0.5
    s_, dp = torch.rand(B, C, device="cuda"), mk(B, C)
    coef = torch.rand(3, C, device="cuda")
    res = {}
    try:
        for mode in (0, 8):
            unetca_b200._lib.load().unetca_set_tuning(3, mode)
            parts = parts_buf(B, 4096)
            n = cint()
            call("unetca_se_bn_bwd_reduce_pool", BF16, ptr(sg), 2 * C, ptr(dpl), C, ptr(pos), ptr(y), C, B, H, W, C, ptr(sc), ptr(sh),
                 ptr(mean), ptr(parts), ctypes.byref(n), stream())
            sums = parts[: B * n.value * 2 * C].view(B, n.value, 2, C).double().sum(1)
            guard = torch.full((B * H * W * C + 8192,), 3.0, dtype=torch.bfloat16, device="cuda")
            dy = guard[: B * H * W * C].view(B, H, W, C)
            dy.fill_(float("nan"))
            call("unetca_bn_bwd_apply_pool", BF16, ptr(sg), 2 * C, ptr(dpl), C, ptr(pos), ptr(y), C, ptr(dy), C, B, H, W, C, ptr(sc),
                 ptr(sh), ptr(mean), ptr(invstd), ptr(s_), ptr(dp), ptr(coef), stream())
            assert torch.all(guard[B * H * W * C:] == 3.0) and not torch.isnan(dy.float()).any()
            res[mode] = (sums, dy.clone())
    finally:
        unetca_b200._lib.load().unetca_set_tuning(3, DEFAULT_APPLY_STREAM)
    assert torch.equal(res[0][1].view(torch.int16), res[8][1].view(torch.int16))
    assert relerr(res[8][0], res[0][0]) < 1e-5


@pytest.mark.parametrize("B,C,H,W,pool", [(2, 64, 8, 128, True), (3, 64, 6, 70, True), (2, 128, 12, 16, True), (1, 256, 4, 34, True),
                                          (1, 64, 2, 2, True), (2, 64, 9, 37, False), (2, 1024, 2, 2, False), (3, 128, 40, 24, False)])
def test_se_scale_pool_stream_equals_register_kernel(B, C, H, W, pool):
    """Block output relu(bn(y)) * s (+ MaxPool2d(2) value and window position, UCA:72, 106-109) through the shared-memory
    streams: bit-identical to the register kernels, also when the output is a channel slice of a wider buffer."""
    rs = np.random.RandomState(23)
    mk = lambda *sh: torch.from_numpy(rs.standard_normal(sh).astype(np.float32)).cuda()
    y = mk(B, H, W, C).bfloat16()
    y[0, 0, 0, :8] = float("nan")                                     # NaN propagates through max-pool like torch
    sc, sh = torch.rand(C, device="cuda") + 0.5, mk(C) * 0.2
    s_ = torch.rand(B, C, device="cuda")
    res = {}
    try:
        for mode in (0, 8):
            unetca_b200._lib.load().unetca_set_tuning(3, mode)
            wide = torch.full((B, H, W, 2 * C if pool else C), 5.0, dtype=torch.bfloat16, device="cuda")
            out = wide[..., :C]
            pooled = torch.full((B, H // 2, W // 2, C), 5.0, dtype=torch.bfloat16, device="cuda") if pool else None
            pos = torch.full((B, H // 2, W // 2, C), 9, dtype=torch.uint8, device="cuda") if pool else None
            call("unetca_se_scale_pool", BF16, ptr(y), C, ptr(out), wide.shape[-1], ptr(pooled), C, ptr(pos), B, H, W, C, ptr(sc), ptr(sh),
                 ptr(s_), stream())
            if pool:
                assert torch.all(wide[..., C:] == 5.0)                 # the other half of the concat buffer is untouched
            res[mode] = (wide.clone(), pooled, pos)
    finally:
        unetca_b200._lib.load().unetca_set_tuning(3, DEFAULT_APPLY_STREAM)
    assert torch.equal(res[0][0].view(torch.int16), res[8][0].view(torch.int16))
    if pool:
        assert torch.equal(res[0][1].view(torch.int16), res[8][1].view(torch.int16))
        assert torch.equal(res[0][2], res[8][2])


@pytest.mark.parametrize("B,HW,nc", [(2, 128, 2), (3, 128 * 37, 2), (2, 512 * 512, 2), (2, 1024, 1)])
def test_outc_stream_equals_register_kernels(B, HW, nc):
    """outc (UCA:125, 162) forward and backward through the shared-memory streams against the register kernels:
    logits and dx bit-identical, dW / db equal up to fp32 summation order."""
    C = 64
    rs = np.random.RandomState(27)
    mk = lambda *sh: torch.from_numpy(rs.standard_normal(sh).astype(np.float32)).cuda()
    x = mk(B, HW, C).bfloat16()
    w, bias = mk(nc, C) * 0.2, mk(nc)
    g = mk(B, nc, HW)
    gscale = torch.tensor([0.37], device="cuda")
    res = {}
    try:
        for mode in (0, 8):
            unetca_b200._lib.load().unetca_set_tuning(3, mode)
            logits = torch.full((B, nc, HW), float("nan"), device="cuda")
            call("unetca_outc_fwd", BF16, ptr(x), C, C, ptr(w), ptr(bias), nc, ptr(logits), B, HW, stream())
            dx = torch.full((B, HW, C), float("nan"), dtype=torch.bfloat16, device="cuda")
            parts = torch.empty(unetca_b200._lib.load().unetca_max_parts(B) * (8 * C + 8), device="cuda")
            dw, db = torch.empty(nc, C, device="cuda"), torch.empty(nc, device="cuda")
            call("unetca_outc_bwd", BF16, ptr(g), ptr(gscale), ptr(x), C, ptr(dx), C, C, ptr(w), nc, B, HW, ptr(parts), ptr(dw), ptr(db),
                 stream())
            res[mode] = (logits, dx, dw, db)
    finally:
        unetca_b200._lib.load().unetca_set_tuning(3, DEFAULT_APPLY_STREAM)
    assert torch.equal(res[0][0], res[8][0]) and not torch.isnan(res[8][0]).any()
    assert torch.equal(res[0][1].view(torch.int16), res[8][1].view(torch.int16))
    assert relerr(res[8][2], res[0][2]) < 1e-5 and relerr(res[8][3], res[0][3]) < 1e-5
    ref = torch.einsum("bpc,oc->bop", x.float(), w) + bias.view(1, nc, 1)
    assert relerr(res[8][0], ref) < 1e-5


@pytest.mark.parametrize("B,H,W,use_se", [(2, 16, 32, True), (3, 64, 64, True), (2, 16, 32, False), (5, 32, 48, True)])
def test_fused_output_head_matches_separate_passes(B, H, W, use_se):
    """unetca_se_scale_outc_fwd / unetca_outc_bn_bwd_reduce / unetca_outc_bn_bwd_apply against the passes they replace
    (se_scale_pool -> outc_fwd; outc_bwd -> se_bn_bwd_reduce / bn_bwd_reduce -> bn_bwd_apply): the block output and its
    gradient are never stored, the results are the same (UCA:86-94 of conv4, UCA:162 and their autograd)."""
    dt, C, nc, HW = BF16, 64, 2, H * W
    rs = np.random.RandomState(B * 7 + H)
    call("unetca_set_conv_impl", 0)
    y2 = to_nhwc(torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32)), dt)
    f = lambda *shape, sc=1.0, off=0.0: torch.from_numpy((off + sc * rs.standard_normal(shape)).astype(np.float32)).cuda()
    scale, shift, mean = f(C, sc=0.2, off=1.0), f(C, sc=0.3), f(C, sc=0.2)
    invstd, coef = f(C, sc=0.1, off=1.0), f(3, C, sc=0.1, off=0.5)
    s = torch.sigmoid(f(B, C)) if use_se else None
    dp = f(B, C, sc=0.5) if use_se else None
    w, bias = f(nc, C, sc=0.2), f(nc, sc=0.1)
    g = f(B, nc, H, W, sc=1.0)
    gscale = torch.tensor([1.0 / (B * HW)], device="cuda")
    # ---- forward: separate
    out = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    call("unetca_se_scale_pool", dt, ptr(y2), C, ptr(out), C, None, 0, None, B, H, W, C, ptr(scale), ptr(shift), ptr(s), stream())
    lg_ref = torch.empty(B, nc, H, W, device="cuda")
    call("unetca_outc_fwd", dt, ptr(out), C, C, ptr(w), ptr(bias), nc, ptr(lg_ref), B, HW, stream())
    lg = torch.full((B, nc, H, W), float("nan"), device="cuda")
    call("unetca_se_scale_outc_fwd", dt, ptr(y2), C, B, HW, C, ptr(scale), ptr(shift), ptr(s), ptr(w), ptr(bias), nc, ptr(lg), stream())
    # (the fused head keeps the block output in fp32 registers where the separate passes stored it as bf16)
    assert relerr(lg.cpu(), lg_ref.cpu()) < 4e-3
    # ---- backward: separate
    parts = parts_buf(B)
    dh = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    dw_ref, db_ref = torch.empty(nc, C, device="cuda"), torch.empty(nc, device="cuda")
    call("unetca_outc_bwd", dt, ptr(g), ptr(gscale), ptr(out), C, ptr(dh), C, C, ptr(w), nc, B, HW, ptr(parts), ptr(dw_ref), ptr(db_ref),
         stream())
    n = cint()
    if use_se:
        call("unetca_se_bn_bwd_reduce", dt, ptr(dh), C, ptr(y2), C, B, HW, C, ptr(scale), ptr(shift), ptr(mean), ptr(parts), ctypes.byref(n),
             stream())
    else:
        call("unetca_bn_bwd_reduce", dt, ptr(dh), C, ptr(y2), C, B, HW, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd), None, None,
             ptr(parts), ctypes.byref(n), stream())
    rows = n.value * (B if use_se else 1)
    sums_ref = parts[: rows * 2 * C].view(rows, 2, C).double().sum(0).cpu()
    dy_ref = torch.empty(B, H, W, C, dtype=TDT[dt], device="cuda")
    call("unetca_bn_bwd_apply", dt, ptr(dh), C, ptr(y2), C, ptr(dy_ref), C, B, HW, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd), ptr(s),
         ptr(dp), ptr(coef), stream())
    # ---- backward: fused
    parts2 = parts_buf(B)
    ws = torch.empty(1 << 20, device="cuda")
    dw, db = torch.empty(nc, C, device="cuda"), torch.empty(nc, device="cuda")
    n2 = cint()
    call("unetca_outc_bn_bwd_reduce", dt, ptr(g), ptr(gscale), ptr(w), nc, ptr(y2), C, B, HW, C, ptr(scale), ptr(shift), ptr(mean), ptr(s),
         ptr(parts2), ctypes.byref(n2), ptr(ws), ws.numel(), ptr(dw), ptr(db), stream())
    sums = parts2[: n2.value * B * 2 * C].view(n2.value * B, 2, C).double().sum(0).cpu()
    assert relerr(sums, sums_ref) < 5e-3
    assert relerr(dw.cpu(), dw_ref.cpu()) < 5e-3 and relerr(db.cpu(), db_ref.cpu()) < 1e-5
    dy = torch.full((B, H, W, C), float("nan"), dtype=TDT[dt], device="cuda")
    call("unetca_outc_bn_bwd_apply", dt, ptr(g), ptr(gscale), ptr(w), nc, ptr(y2), C, ptr(dy), C, B, HW, C, ptr(scale), ptr(shift), ptr(mean),
         ptr(invstd), ptr(s), ptr(dp), ptr(coef), stream())
    assert relerr(dy.float().cpu(), dy_ref.float().cpu()) < 1e-2          # dh unrounded vs bf16-rounded


@pytest.mark.parametrize("B,C,O,H,W,split", [(2, 64, 128, 40, 24, 64), (1, 128, 256, 33, 17, 128), (2, 256, 512, 16, 16, 256)])
def test_conv3x3_two_destination_epilogue(B, C, O, H, W, split):
    """unetca_conv3x3_fwd_split == unetca_conv3x3_fwd with the output channels delivered as two dense tensors (the dgrad of a
    decoder block's first conv: d skip / d upsampled, autograd of torch.cat at UCA:140), statistics unchanged."""
    dt = BF16
    rs = np.random.RandomState(C + O)
    call("unetca_set_conv_impl", 0)
    x = to_nhwc(torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32)), dt)
    w = torch.from_numpy((rs.standard_normal((O, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32)).cuda()
    wf = torch.empty(O, 9 * C, dtype=TDT[dt], device="cuda")
    call("unetca_pack_conv3x3_weight", dt, ptr(w), ptr(wf), 9 * C, None, O, C, stream())
    y = torch.empty(B, H, W, O, dtype=TDT[dt], device="cuda")
    parts, n = parts_buf(B), cint()
    call("unetca_conv3x3_fwd", dt, ptr(x), C, ptr(wf), 9 * C, ptr(y), O, B, H, W, C, O, ptr(parts), ctypes.byref(n), stream())
    st_ref = parts[: n.value * 2 * O].clone()
    y1 = torch.full((B, H, W, split), float("nan"), dtype=TDT[dt], device="cuda")
    y2 = torch.full((B, H, W, O - split), float("nan"), dtype=TDT[dt], device="cuda")
    parts2, n2 = parts_buf(B), cint()
    call("unetca_conv3x3_fwd_split", dt, ptr(x), C, ptr(wf), 9 * C, ptr(y1), split, ptr(y2), O - split, split, B, H, W, C, O,
         ptr(parts2), ctypes.byref(n2), stream())
    assert torch.equal(y1, y[..., :split]) and torch.equal(y2, y[..., split:])
    assert n2.value == n.value and torch.equal(parts2[: n.value * 2 * O], st_ref)


@pytest.mark.parametrize("B,C,O,H,W", [(2, 128, 64, 40, 24), (1, 128, 64, 33, 17), (2, 256, 128, 24, 40), (1, 512, 256, 17, 33),
                                       (2, 1024, 512, 16, 16), (3, 128, 64, 64, 64), (2, 256, 128, 64, 64)])
def test_conv3x3_two_source_operand(B, C, O, H, W):
    """unetca_conv3x3_fwd_cat / _wgrad_cat == the one-tensor entries on torch.cat([skip, up], 1) (UCA:140): the decoder's first
    conv reading its two input halves from two dense tensors.  Output, statistics and filter gradient bit-identical; also the
    eval-mode BatchNorm + ReLU (+ SE squeeze) epilogue forms."""
    dt = BF16
    rs = np.random.RandomState(C + O + H)
    call("unetca_set_conv_impl", 0)
    C1 = C // 2
    x = to_nhwc(torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32)), dt)
    x1, x2 = x[..., :C1].contiguous(), x[..., C1:].contiguous()
    w = torch.from_numpy((rs.standard_normal((O, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32)).cuda()
    wf = torch.empty(O, 9 * C, dtype=TDT[dt], device="cuda")
    call("unetca_pack_conv3x3_weight", dt, ptr(w), ptr(wf), 9 * C, None, O, C, stream())
    y = torch.empty(B, H, W, O, dtype=TDT[dt], device="cuda")
    y2 = torch.full((B, H, W, O), float("nan"), dtype=TDT[dt], device="cuda")
    parts, n = parts_buf(B), cint()
    parts2, n2 = parts_buf(B), cint()
    if O == 64:
        wk = torch.empty(9 * C, 64, dtype=TDT[dt], device="cuda")
        call("unetca_pack_conv3x3_kw", dt, ptr(wf), 9 * C, ptr(wk), C, stream())
        call("unetca_conv3x3_fwd_kw", dt, ptr(x), C, ptr(wk), ptr(y), O, B, H, W, C, ptr(parts), ctypes.byref(n), stream())
        wsel, layout = wk, 2
    else:
        call("unetca_conv3x3_fwd", dt, ptr(x), C, ptr(wf), 9 * C, ptr(y), O, B, H, W, C, O, ptr(parts), ctypes.byref(n), stream())
        wsel, layout = wf, 0
    call("unetca_conv3x3_fwd_cat", dt, ptr(x1), C1, ptr(x2), C - C1, C1, ptr(wsel), ptr(y2), O, B, H, W, C, O, ptr(parts2), None, None,
         None, ctypes.byref(n2), stream())
    assert torch.equal(y, y2)
    assert n.value == n2.value and torch.equal(parts[: n.value * 2 * O], parts2[: n.value * 2 * O])
    # eval-mode epilogue
    scale = torch.from_numpy(rs.uniform(0.5, 1.5, O).astype(np.float32)).cuda()
    shift = torch.from_numpy(rs.standard_normal(O).astype(np.float32) * 0.1).cuda()
    nsm = unetca_b200._lib.load().unetca_num_sms()
    sq = torch.zeros(B * nsm * O, device="cuda") if O != 64 else None
    sq2 = torch.zeros(B * nsm * O, device="cuda") if O != 64 else None
    a = torch.empty_like(y)
    a2 = torch.full_like(y, float("nan"))
    call("unetca_conv3x3_bnrelu_fwd", dt, ptr(x), C, ptr(wsel), layout, ptr(a), O, B, H, W, C, O, ptr(scale), ptr(shift),
         ptr(sq) if sq is not None else None, ctypes.byref(n), stream())
    call("unetca_conv3x3_fwd_cat", dt, ptr(x1), C1, ptr(x2), C - C1, C1, ptr(wsel), ptr(a2), O, B, H, W, C, O, None, ptr(scale), ptr(shift),
         ptr(sq2) if sq2 is not None else None, ctypes.byref(n2), stream())
    assert torch.equal(a, a2)
    if sq is not None:
        assert n.value == n2.value and torch.equal(sq, sq2)
    # filter gradient
    dy = to_nhwc(torch.from_numpy(rs.standard_normal((B, O, H, W)).astype(np.float32)), dt)
    ws = torch.empty(16 * 1024 * 1024, device="cuda")
    dw, dw2 = torch.empty(O, C, 3, 3, device="cuda"), torch.full((O, C, 3, 3), float("nan"), device="cuda")
    call("unetca_conv3x3_wgrad", dt, ptr(dy), O, ptr(x), C, ptr(ws), ws.numel(), B, H, W, C, O, ptr(dw), stream())
    call("unetca_conv3x3_wgrad_cat", dt, ptr(dy), O, ptr(x1), C1, ptr(x2), C - C1, C1, ptr(ws), ws.numel(), B, H, W, C, O, ptr(dw2),
         stream())
    assert torch.equal(dw, dw2)
