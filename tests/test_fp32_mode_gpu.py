"""GPU: the fp32 mode (BASELINE.json north_star: "loss and logits within ... 1e-4 for the fp32 mode").

fp32 NDHWC activations; every convolution is the sum of bf16 split-plane products on the tcgen05 kernels
(csrc/fp32_mode.cu, ConvTileParams::out_f32). Checked here through the C ABI:
  * the conv trio against an fp64 evaluation of the same convolution on the UNROUNDED fp32 operands,
  * the fp32 BatchNorm / pooling kernels against torch,
  * whole encoders and full SimCLR+DualVar / MoCo+DualVar steps against the oracle in fp32 (TF32 off): losses and
    logits within 1e-4 relative, with the oracle's own distance to an fp64 run of itself printed as the yardstick.
"""
import copy
import ctypes
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda:0"
ARGS = SimpleNamespace(shufflerank_theta=0.05)
TOL = 1e-4   # the north star's fp32-mode tolerance


@pytest.fixture(autouse=True)
def _fp32_mode():
    from dualvar_b200 import engine as E
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    E.set_precision("fp32", planes=3)
    yield
    E.set_precision("bf16")


def _seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _relmax(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def _to_ndhwc_f32(x, Cp):
    N, C = x.shape[:2]
    out = torch.zeros((N,) + tuple(x.shape[2:]) + (Cp,), dtype=torch.float32, device=x.device)
    out[..., :C] = x.permute(0, 2, 3, 4, 1)
    return out.contiguous()


def _planes(x_nd, K):
    from dualvar_b200 import _lib
    p = torch.empty((K,) + tuple(x_nd.shape), dtype=torch.bfloat16, device=x_nd.device)
    _lib.call("dv_f32_split", _lib.ptr(x_nd), _lib.ptr(p), p.stride(0), K, x_nd.numel(), _lib.stream_ptr())
    return p


CONV_CASES = [
    ("1x1x1 64->64", 2, 4, 16, 16, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("spatial 64->144", 2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("temporal 144->64", 2, 4, 14, 14, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("spatial s2 64->230", 2, 4, 28, 28, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("temporal s2 230->128", 2, 8, 14, 14, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("down 1x1 s(1,2,2) 64->42", 2, 4, 28, 28, 64, 42, (1, 1, 1), (1, 2, 2), (0, 0, 0)),
    ("3x3x3 s2 64->128", 2, 8, 14, 14, 64, 128, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    ("spatial 128->288 (2 n-tiles)", 2, 4, 14, 14, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("odd extents 5x9x11", 3, 5, 9, 11, 40, 72, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("two-region tiling 64->144 24x32", 2, 2, 24, 32, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("pairs 64->64 many tiles", 24, 8, 32, 32, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
]


@pytest.mark.parametrize("K,tol", [(3, 5e-6), (2, 4e-5)])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_trio_fp32_mode_matches_fp64(case, K, tol):
    """Sum of split-plane products vs the fp64 convolution of the unrounded fp32 operands (max error relative to the
    output range). 3 planes: fp32-level (5e-6: fp32 accumulation over up to 2592 terms, measured 3.4e-6); 2 planes:
    16 mantissa bits (4e-5)."""
    from dualvar_b200 import _lib, engine as E
    import kernel_handles as KK
    E.set_precision("fp32", planes=K)
    name, N, T, H, W, Cin, Cout, k, s, p = case
    g = KK.make_geom(N, T, H, W, Cin, Cout, k, s, p)
    gen = torch.Generator(device=dev).manual_seed(abs(hash(name)) % (2 ** 31))
    x = torch.randn(N, Cin, T, H, W, device=dev, generator=gen)
    w = torch.randn(Cout, Cin, *k, device=dev, generator=gen) / (Cin * k[0] * k[1] * k[2]) ** 0.5
    bias = torch.randn(Cout, device=dev, generator=gen)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    yr = F.conv3d(xr, wr, bias.double(), s, p)
    dy = torch.randn(yr.shape, device=dev, generator=gen)
    yr.backward(dy.double())

    call, ptr, st = _lib.call, _lib.ptr, _lib.stream_ptr
    xp = _planes(_to_ndhwc_f32(x, g.Cin_p), K)
    conv = torch.nn.Conv3d(Cin, Cout, k, s, p, bias=False).to(dev)
    with torch.no_grad():
        conv.weight.copy_(w)
    wp = E.packed_weight_planes(conv)
    bias_p = torch.zeros(g.Cout_p, device=dev)
    bias_p[:Cout] = bias
    y = torch.full((N, g.To, g.Ho, g.Wo, g.Cout_p), float("nan"), dtype=torch.float32, device=dev)   # the first product overwrites
    for n, (i, j) in enumerate(E._terms()):
        call("dv_conv3d_fprop_f32acc", ptr(xp[i]), ptr(wp[j][0]), ptr(y), ptr(bias_p) if n == 0 else None,
             ctypes.byref(g), 1 if n else 0, st())
    assert _relmax(y[..., :Cout].permute(0, 4, 1, 2, 3), yr.detach()) < tol
    if g.Cout_p > Cout:
        assert bool((y[..., Cout:] == 0).all())
    # the same products as extra taps of ONE launch (the engine's path for stride-1 layers): same bound
    wf_all, wt_all = E.packed_weight_planes_all(conv)
    ym = torch.full_like(y, float("nan"))
    n_parity = int(np.prod([min(a, b) for a, b in zip(k, s)]))
    if n_parity * K <= 12:      # one tensor map per (stride-parity class, plane)
        mstats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
        call("dv_conv3d_fprop_f32planes", ptr(xp), xp.stride(0), K, ptr(wf_all), ptr(ym), ptr(mstats), ptr(bias_p),
             ctypes.byref(g), st())
        assert _relmax(ym[..., :Cout].permute(0, 4, 1, 2, 3), yr.detach()) < tol
        assert _relmax(ym, y) < tol
        if g.Cout_p > Cout:
            assert bool((ym[..., Cout:] == 0).all())
        # the epilogue's batch statistics (fp32 partials per CTA, double across CTAs) of the stored output
        ysm = ym.reshape(-1, g.Cout_p).double()
        scale = ysm.abs().sum(0).clamp_min(1e-30)
        assert float(((mstats[:g.Cout_p] - ysm.sum(0)).abs() / scale).max()) < 2e-6
        assert float(((mstats[g.Cout_p:] - (ysm * ysm).sum(0)).abs() / (ysm * ysm).sum(0).clamp_min(1e-30)).max()) < 2e-6
    else:
        with pytest.raises(_lib.DualVarNativeError, match="views"):
            call("dv_conv3d_fprop_f32planes", ptr(xp), xp.stride(0), K, ptr(wf_all), ptr(ym), None, ptr(bias_p),
                 ctypes.byref(g), st())
    # batch statistics of the fp32 output
    stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
    call("dv_f32_colstats", ptr(y), ptr(stats), y.numel() // g.Cout_p, g.Cout_p, st())
    ys = y.reshape(-1, g.Cout_p).double()
    torch.testing.assert_close(stats[:g.Cout_p], ys.sum(0), rtol=1e-9, atol=1e-6)
    torch.testing.assert_close(stats[g.Cout_p:], (ys * ys).sum(0), rtol=1e-9, atol=1e-6)

    dyp = _planes(_to_ndhwc_f32(dy, g.Cout_p), K)
    dx = torch.full((N, T, H, W, g.Cin_p), float("nan"), dtype=torch.float32, device=dev)
    for n, (i, j) in enumerate(E._terms()):
        call("dv_conv3d_dgrad_f32acc", ptr(dyp[i]), ptr(wp[j][1]), ptr(dx), ctypes.byref(g), 1 if n else 0, st())
    assert _relmax(dx[..., :Cin].permute(0, 4, 1, 2, 3), xr.grad) < tol
    dxm = torch.full_like(dx, float("nan"))      # merged launch (one per stride-parity class), any stride
    call("dv_conv3d_dgrad_f32planes", ptr(dyp), dyp.stride(0), K, ptr(wt_all), ptr(dxm), ctypes.byref(g), st())
    assert _relmax(dxm[..., :Cin].permute(0, 4, 1, 2, 3), xr.grad) < tol
    assert _relmax(dxm, dx) < tol
    dwp = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.float32, device=dev)
    for n, (i, j) in enumerate(E._terms()):
        call("dv_conv3d_wgrad_bf16" if n == 0 else "dv_conv3d_wgrad_bf16_acc", ptr(xp[i]), ptr(dyp[j]), ptr(dwp),
             ctypes.byref(g), st())
    dw = KK.unpack_conv_wgrad(dwp, g)
    assert _relmax(dw, wr.grad) < tol
    dwm = torch.full_like(dwp, float("nan"))     # all products in one launch, one reduction into the packed buffer
    call("dv_conv3d_wgrad_f32planes", ptr(xp), ptr(dyp), K, ptr(dwm), ctypes.byref(g), st())
    assert _relmax(KK.unpack_conv_wgrad(dwm, g), wr.grad) < tol
    assert _relmax(dwm, dwp) < tol


def test_split_planes_are_exact():
    """x == p0 + p1 + p2 to the last bit for normal fp32 values (24 = 3 x 8 mantissa bits), residuals exact."""
    x = torch.randn(1 << 16, device=dev) * torch.logspace(-3, 3, 1 << 16, device=dev)
    p = _planes(x.view(-1, 8), 3).float()
    assert torch.equal(p.sum(0).view(-1), x)
    p2 = _planes(x.view(-1, 8), 2).float()
    assert ((p2.sum(0).view(-1) - x).abs() <= x.abs() * 2.0 ** -16).all()


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("mode", ["plain", "res", "two"])
def test_f32_batchnorm_unit_matches_torch(mode, relu):
    """conv output -> training-mode BN (+ residual / second BN branch) (+ ReLU), forward and backward, fp32 kernels
    vs torch autograd in fp64."""
    from dualvar_b200 import engine as E
    import torch.nn as nn
    _seed(3)
    N, T, H, W, C = 3, 4, 10, 12, 72
    x = torch.randn(N, C, T, H, W, device=dev)
    conv1 = nn.Conv3d(C, C, (1, 3, 3), 1, (0, 1, 1), bias=False).to(dev)
    conv2 = nn.Conv3d(C, C, 1, 1, 0, bias=False).to(dev)
    bn1, bn2 = nn.BatchNorm3d(C).to(dev), nn.BatchNorm3d(C).to(dev)
    with torch.no_grad():
        for bn in (bn1, bn2):
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    mods = [conv1, conv2, bn1, bn2]
    ref = [copy.deepcopy(m).double() for m in mods]

    def torch_fwd(xd):
        c1, c2, b1, b2 = ref
        h = F.relu(b1(c1(xd)))            # the residual / input of the unit under test
        o = b2(c2(h))
        if mode == "res":
            o = o + h
        elif mode == "two":
            o = o + b1(c1(xd))            # second BN branch (fresh statistics of the same conv)
        return F.relu(o) if relu else o

    xd = x.double()
    out_r = torch_fwd(xd)
    dout = torch.randn_like(out_r)
    out_r.backward(dout)

    ctx = E.Context(training=True)
    xin = E.Act(_to_ndhwc_f32(x, 72), C, needs_grad=False, planes=_planes(_to_ndhwc_f32(x, 72), 3))
    h = E.activate(ctx, E.conv_stats(ctx, xin, conv1, bn1))
    main = E.conv_stats(ctx, h, conv2, bn2)
    if mode == "res":
        o = E.activate(ctx, main, res=h, relu=relu)
    elif mode == "two":
        bn1b = copy.deepcopy(bn1)
        o = E.activate(ctx, main, r2=E.conv_stats(ctx, xin, conv1, bn1b), relu=relu)
    else:
        o = E.activate(ctx, main, relu=relu)
    assert _relmax(o.data.permute(0, 4, 1, 2, 3), out_r.detach()) < 2e-5
    o.grad = _to_ndhwc_f32(dout.float(), 72)
    E.run_backward(ctx)
    got = {id(p): gval for p, gval in ((p, ctx.param_grads.get(id(p))) for m in mods for p in m.parameters())}
    for m, mr in zip(mods, ref):
        for (n, p), (_, pr) in zip(m.named_parameters(), mr.named_parameters()):
            gp = got[id(p)]
            assert gp is not None, n
            if mode == "two" and m is bn1:
                continue          # bn1 is used by two branches in the torch graph, one of them via the copy here
            assert _rel(gp, pr.grad) < 5e-4, (type(m).__name__, n, _rel(gp, pr.grad))


def test_f32_pooling_matches_torch():
    from dualvar_b200 import engine as E
    _seed(4)
    x = torch.randn(2, 24, 6, 12, 12, device=dev).relu()     # ties at 0 exercise the first-maximum rule
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool3d(xr, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    pr = F.adaptive_avg_pool3d(yr, 1).flatten(1)
    dp = torch.randn_like(pr)
    pr.backward(dp)
    ctx = E.Context(training=True)
    xa = E.Act(_to_ndhwc_f32(x, 24), 24)
    ya = E.max_pool(ctx, xa, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    assert torch.equal(ya.data.permute(0, 4, 1, 2, 3), yr.detach())
    assert torch.equal(ya.planes.float().sum(0), ya.data)
    pooled = E.global_pool(ctx, ya)
    torch.testing.assert_close(pooled, pr.detach(), rtol=1e-6, atol=1e-6)
    E.global_pool_backward(ya, dp)
    E.run_backward(ctx)
    torch.testing.assert_close(xa.grad.permute(0, 4, 1, 2, 3), xr.grad, rtol=1e-6, atol=1e-7)
    back = E.to_ncdhw(xa)
    assert torch.equal(back, x)


def _try64(fn):
    """fp64 run of the oracle as the yardstick for both fp32 paths; None if the oracle cannot run in double."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001 - the yardstick is informative, the assertions do not depend on it
        print(f"fp64 yardstick unavailable: {type(e).__name__}: {e}")
        return None


def _grad_report(tag, ref, prod, ref64):
    ours, ours64, yard = [], [], []
    for (n, pr), (_, pp) in zip(ref.named_parameters(), prod.named_parameters()):
        if pr.grad is None:
            assert pp.grad is None, n
            continue
        assert pp.grad is not None and torch.isfinite(pp.grad).all(), n
        ours.append(_rel(pp.grad, pr.grad))
    if ref64 is not None:
        for (n, pr), (_, pp), (_, p64) in zip(ref.named_parameters(), prod.named_parameters(), ref64.named_parameters()):
            if pr.grad is not None and p64.grad is not None:
                ours64.append(_rel(pp.grad, p64.grad)); yard.append(_rel(pr.grad, p64.grad))
    med = lambda v: sorted(v)[len(v) // 2] if v else float("nan")  # noqa: E731
    print(f"{tag}: grad rel L2 ours vs oracle fp32 median {med(ours):.2e} max {max(ours):.2e}; vs fp64: ours "
          f"{med(ours64):.2e}, oracle fp32 {med(yard):.2e}")
    # fp32-level agreement: the median parameter gradient within 1e-3 (or 10x the oracle's own fp32 rounding noise)
    bound = max(1e-3, 10 * med(yard)) if yard else 1e-3
    assert med(ours) <= bound, (med(ours), bound)
    assert max(ours) <= 50 * bound, (max(ours), bound)


@pytest.mark.parametrize("net,shape", [("r21d", (4, 3, 8, 64, 64)), ("r3d", (4, 3, 8, 64, 64)),
                                       ("c3d", (4, 3, 8, 64, 64)), ("r2d3d18", (4, 3, 4, 96, 96)),
                                       ("s3d", (3, 3, 32, 128, 128)), ("s3dg", (3, 3, 32, 128, 128))])
def test_backbone_fp32_mode_matches_oracle(net, shape):
    """select_backbone(net) forward + backward in the fp32 mode vs the oracle backbone in fp32 (and in fp64)."""
    from dualvar_b200 import backbones as PB
    from oracle import backbones as OB
    _seed(1)
    ref, _ = OB.select_backbone(net)
    ref = ref.to(dev).train()
    prod, _ = PB.select_backbone(net)
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev).train()
    x = torch.randn(*shape, device=dev)
    yr, yp = ref(x), prod(x)
    ref64 = _try64(lambda: copy.deepcopy(ref).double())
    y64 = _try64(lambda: ref64(x.double())) if ref64 is not None else None
    if y64 is not None:
        print(f"{net}: forward max err / range vs fp64: ours {_relmax(yp, y64):.2e}, oracle fp32 {_relmax(yr, y64):.2e}")
    print(f"{net}: forward ours vs oracle fp32: {_relmax(yp, yr):.2e}")
    assert yp.shape == yr.shape and yp.dtype == yr.dtype
    # within 1e-4 of the oracle - or, where the oracle's own fp32 run is further than that from the exact result
    # (S3D: N(0, 0.01) weights, BatchNorm over few positions in the last stages), at least as close to fp64 as it is
    assert _relmax(yp, yr) < TOL or (y64 is not None and _relmax(yp, y64) <= 2 * _relmax(yr, y64))
    w = torch.randn_like(yr)
    (yr * w).sum().backward(); (yp * w).sum().backward()
    if y64 is not None:
        (y64 * w.double()).sum().backward()
    _grad_report(net, ref, prod, ref64 if y64 is not None else None)
    for (n, br), (_, bp) in zip(ref.named_buffers(), prod.named_buffers()):
        if br.dtype.is_floating_point:
            assert _rel(bp, br) < 1e-4, n


@pytest.mark.parametrize("net,shape", [("r21d", (4, 3, 8, 64, 64)), ("r3d", (4, 3, 8, 64, 64))])
def test_merged_and_per_product_engine_paths_agree(net, shape):
    """engine.F32_MERGE (all plane products of a convolution in one launch, statistics from the epilogue) against one
    launch per product + dv_f32_colstats on the same backbone: output, parameter gradients and running statistics."""
    from dualvar_b200 import backbones as PB, engine as E, _lib
    _seed(3)
    a, _ = PB.select_backbone(net)
    a = a.to(dev).train()
    b = copy.deepcopy(a)
    x = torch.randn(*shape, device=dev)
    res = []
    for m, merge in ((a, True), (b, False)):
        old = E.F32_MERGE
        E.F32_MERGE = merge
        try:
            n0 = _lib.load().dv_launch_count()
            y = m(x)
            (y * torch.linspace(-1, 1, y.numel(), device=dev).view_as(y)).sum().backward()
            res.append((y.detach(), _lib.load().dv_launch_count() - n0))
        finally:
            E.F32_MERGE = old
    (ya, la), (yb, lb) = res
    assert la < 0.7 * lb, (la, lb)            # the merged path really ran (well under half of the launches)
    assert _relmax(ya, yb) < 5e-6
    worst = max(_rel(pa.grad, pb.grad) for pa, pb in zip(a.parameters(), b.parameters()) if pa.grad is not None)
    print(f"{net}: merged vs per-product: output {_relmax(ya, yb):.2e}, worst gradient rel L2 {worst:.2e}, launches {la} vs {lb}")
    assert worst < 2e-4, worst
    for (n, ba), (_, bb) in zip(a.named_buffers(), b.named_buffers()):
        if ba.dtype.is_floating_point:
            assert _relmax(ba, bb) < 1e-5, n


def _model_pair(kind, net):
    from dualvar_b200 import models as PM
    from oracle import models as OM
    _seed(0)
    if kind == "simclr":
        a = (net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS)
        ref, prod = OM.SimCLR_TimeSeriesV4(*a), PM.SimCLR_TimeSeriesV4(*a)
    else:
        a = (net, 128, 256, 0.999, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS)
        ref, prod = OM.MoCo_TimeSeriesV4(*a), PM.MoCo_TimeSeriesV4(*a)
        # a freshly built MoCo has key encoder == query encoder, so k == q and the clip loss is ~1e-3 (a difference of
        # two ~14.3 logits: ill-conditioned in any fp32 evaluation). Move the key encoder off the query encoder, as
        # training does after the first momentum steps.
        with torch.no_grad():
            for n, prm in ref.named_parameters():
                if n.startswith(("encoder_k", "series_proj_head_k")):
                    prm.mul_(1.0 + 0.1 * torch.randn_like(prm))
    ref = ref.to(dev).train()
    prod.load_state_dict(ref.state_dict())
    return ref, prod.to(dev).train()


@pytest.mark.parametrize("kind,net", [("simclr", "r21d"), ("simclr", "r3d"), ("moco", "r21d")])
def test_dualvar_step_fp32_mode_within_1e4_of_oracle(kind, net):
    """One full pretraining step (3 views, clip + tc + shuffle-rank objectives): every loss within 1e-4 relative of
    the oracle in fp32, logits within 1e-4 of the logit range, gradients at the fp32 noise level."""
    from dualvar_b200 import _lib
    ref, prod = _model_pair(kind, net)
    ref64 = _try64(lambda: copy.deepcopy(ref).double())
    x = torch.randn(8, 3, 3, 8, 64, 64, device=dev)
    n0 = _lib.load().dv_launch_count()
    np.random.seed(11); torch.manual_seed(5); rr = ref(x)
    np.random.seed(11); torch.manual_seed(5); rp = prod(x)

    def run64():
        np.random.seed(11); torch.manual_seed(5)
        return ref64(x.double())
    r64 = _try64(run64) if ref64 is not None else None
    assert list(rr.keys()) == list(rp.keys())
    for k in rr:
        if "labels" in k:
            assert torch.equal(rr[k], rp[k])
        elif "loss" in k:
            e = abs(rp[k].item() - rr[k].item()) / abs(rr[k].item())
            y = f" (oracle fp32 vs fp64 {abs(rr[k].item() - r64[k].item()) / abs(r64[k].item()):.2e})" if r64 else ""
            print(f"{kind}/{net} {k}: ours {rp[k].item():.7f} oracle {rr[k].item():.7f} rel {e:.2e}{y}")
            assert e <= TOL, (k, rp[k].item(), rr[k].item())
        else:
            e = _relmax(rp[k], rr[k])
            y = f" (oracle fp32 vs fp64 {_relmax(rr[k], r64[k]):.2e})" if r64 else ""
            print(f"{kind}/{net} {k}: max err / range {e:.2e}{y}")
            assert e < TOL, (k, e)
    sum(v for k, v in rr.items() if "loss" in k).backward()
    sum(v for k, v in rp.items() if "loss" in k).backward()
    if r64 is not None:
        sum(v for k, v in r64.items() if "loss" in k).backward()
    assert _lib.load().dv_launch_count() - n0 > 300          # the native path really ran (merged plane products: 1 fprop + 1 dgrad + 6 wgrad launches per stride-1 conv)
    _grad_report(f"{kind}/{net}", ref, prod, ref64 if r64 is not None else None)
    for (n, br), (_, bp) in zip(ref.named_buffers(), prod.named_buffers()):
        if br.dtype.is_floating_point and "queue" not in n:
            assert _rel(bp, br) < 1e-4, n


def test_two_plane_mode_step_error_is_reported():
    """2 split planes (3 launches per conv, 16 mantissa bits): the cheaper fp32-mode setting. Its loss error is
    printed and must stay within 10x the tolerance; the 1e-4 claim is made for 3 planes only."""
    from dualvar_b200 import engine as E
    ref, prod = _model_pair("simclr", "r21d")
    E.set_precision("fp32", planes=2)
    x = torch.randn(8, 3, 3, 8, 64, 64, device=dev)
    np.random.seed(11); rr = ref(x)
    np.random.seed(11); rp = prod(x)
    for k in rr:
        if "loss" in k:
            e = abs(rp[k].item() - rr[k].item()) / abs(rr[k].item())
            print(f"2 planes {k}: rel {e:.2e}")
            assert e <= 10 * TOL, (k, e)


def test_fp32_mode_s3dg_step_matches_oracle():
    """S3D-G SimCLR+DualVar step in the fp32 mode at the config-4 geometry (32x128x128; concat slices, self-gating,
    13 max-pools on fp32 tensors). The oracle's own fp32 run of this network sits ~1e-4..1e-3 from its fp64 run, so
    each loss must be within 1e-4 of the oracle OR at least as close to the fp64 value as the oracle's fp32 one (x2)."""
    ref, prod = _model_pair("simclr", "s3dg")
    ref64 = _try64(lambda: copy.deepcopy(ref).double())
    x = torch.randn(3, 3, 3, 32, 128, 128, device=dev)
    np.random.seed(11); rr = ref(x)
    np.random.seed(11); rp = prod(x)

    def run64():
        np.random.seed(11)
        return ref64(x.double())
    r64 = _try64(run64) if ref64 is not None else None
    for k in rr:
        if "labels" in k:
            assert torch.equal(rr[k], rp[k])
        elif "loss" in k:
            e = abs(rp[k].item() - rr[k].item()) / abs(rr[k].item())
            e64 = abs(rp[k].item() - r64[k].item()) / abs(r64[k].item()) if r64 else float("nan")
            y64 = abs(rr[k].item() - r64[k].item()) / abs(r64[k].item()) if r64 else float("nan")
            print(f"simclr/s3dg {k}: ours {rp[k].item():.7f} oracle {rr[k].item():.7f} rel {e:.2e}; vs fp64: ours "
                  f"{e64:.2e}, oracle fp32 {y64:.2e}")
            assert e <= TOL or (r64 is not None and e64 <= 2 * y64), (k, e, e64, y64)
        else:
            e = _relmax(rp[k], rr[k])
            assert e < TOL or (r64 is not None and _relmax(rp[k], r64[k]) <= 2 * _relmax(rr[k], r64[k])), (k, e)
    sum(v for k, v in rr.items() if "loss" in k).backward()
    sum(v for k, v in rp.items() if "loss" in k).backward()
    if r64 is not None:
        sum(v for k, v in r64.items() if "loss" in k).backward()
    _grad_report("simclr/s3dg", ref, prod, ref64 if r64 is not None else None)


def test_fp32_mode_features_give_the_oracles_retrieval_ranking():
    """Config 5 end to end in the fp32 mode: eval-mode R3D features of 96 'train' and 32 'test' clips from the product
    and from the oracle, then the top-k ranking of each (classifier.py:963-983): the fp32-mode features are close
    enough to the oracle's that the retrieved indices agree (top-1 everywhere; deeper ranks except exact near-ties)."""
    from dualvar_b200 import backbones as PB
    from dualvar_b200.retrieval import retrieval_topk
    from oracle import backbones as OB
    _seed(2)
    ref, _ = OB.select_backbone("r3d")
    ref = ref.to(dev).train()
    with torch.no_grad():
        for _ in range(2):                      # give the running statistics some content
            ref(torch.randn(8, 3, 8, 64, 64, device=dev))
    ref.eval()
    prod, _ = PB.select_backbone("r3d")
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev).eval()
    feats = {}
    with torch.no_grad():
        for name, n, seed in (("train", 96, 3), ("test", 32, 4)):
            x = torch.randn(n, 3, 8, 64, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
            fr = torch.cat([ref(x[i:i + 32]).mean((2, 3, 4)) for i in range(0, n, 32)])
            fp = torch.cat([prod.encode(x[i:i + 32].contiguous(), pooled=True) for i in range(0, n, 32)])
            print(f"{name} features: max err / range {_relmax(fp, fr):.2e}")
            assert _relmax(fp, fr) < TOL
            feats[name] = (fr, fp)
    _, top_r = retrieval_topk(feats["test"][0], feats["train"][0])
    _, top_p = retrieval_topk(feats["test"][1], feats["train"][1])
    assert torch.equal(top_r[1], top_p[1])
    agree = (top_r[50] == top_p[50]).float().mean().item()
    print(f"top-50 index agreement {agree:.4f}")
    assert agree > 0.99
