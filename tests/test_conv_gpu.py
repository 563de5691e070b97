"""GPU: tcgen05 implicit-GEMM conv3d fprop / dgrad / wgrad through the C ABI vs torch fp32 conv on
bf16-rounded operands. Tolerance 1e-2 of the output range (bf16 storage of y/dx; wgrad is fp32)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    ("1x1x1 64->64", 2, 4, 16, 16, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("spatial 64->144", 2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("temporal 144->64", 2, 4, 14, 14, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("spatial s2 64->230", 2, 4, 28, 28, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("temporal s2 230->128", 2, 8, 14, 14, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("down 1x1 s(1,2,2) 64->42", 2, 4, 28, 28, 64, 42, (1, 1, 1), (1, 2, 2), (0, 0, 0)),
    ("down 1x1 s(2,1,1) 42->128", 2, 8, 14, 14, 42, 128, (1, 1, 1), (2, 1, 1), (0, 0, 0)),
    ("stem 3->83 7x7 s2", 2, 4, 32, 32, 3, 83, (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    ("r3d stem 3->64 3x7x7", 1, 4, 32, 32, 3, 64, (3, 7, 7), (1, 2, 2), (1, 3, 3)),
    ("3x3x3 64->64", 2, 4, 14, 14, 64, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("3x3x3 s2 64->128", 2, 8, 14, 14, 64, 128, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    ("spatial 128->288 (2 n-tiles)", 2, 4, 14, 14, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("spatial 256->576 7x7", 3, 2, 7, 7, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("temporal 921->512 s2", 3, 4, 7, 7, 921, 512, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("many tiles per CTA 64->64", 24, 8, 32, 32, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("odd extents 5x9x11", 3, 5, 9, 11, 40, 72, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("spatial halo wgrad 64->144 16x16 (2 unit groups)", 2, 4, 16, 16, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("spatial halo wgrad 24->40 24x16", 3, 2, 24, 16, 24, 40, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("two-region tiling 64->144 24x32", 2, 2, 24, 32, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("two-region tiling 3x3x3 40->64 40x16", 2, 3, 40, 16, 40, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    # narrow 1x1x1 convolutions over many input channels (S3D Inception branches): >= 24 MMAs per 64-column tile, i.e. the
    # two-issuer kernel instance, on maps of one or two tiles, 16..64 output columns, partial last K chunk
    ("1x1x1 480->16 4x4x4", 4, 4, 4, 4, 480, 16, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("1x1x1 480->64 4x4x4", 4, 4, 4, 4, 480, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("1x1x1 512->24 4x4x4", 4, 4, 4, 4, 512, 24, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("1x1x1 832->48 2x2x2", 4, 2, 2, 2, 832, 48, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("1x1x1 400->40 8x8x8", 4, 8, 8, 8, 400, 40, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
]


def _rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def test_wgrad_cluster_of_two_unit_groups_matches_torch(monkeypatch):
    """DV_WGRAD_CLUSTER=1: the two unit groups of the 64->144 3x3 weight gradient as a cluster of 2 CTAs in lockstep, each
    fetching half of the dY boxes of a stage with TMA multicast (opt-in, csrc/conv_wgrad.cu wgrad_launch): same result."""
    import kernel_handles as K
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda:0"
    name, N, T, H, W, Cin, Cout, k, s, p = [c for c in CASES if "2 unit groups" in c[0]][0]
    g = K.make_geom(N, T, H, W, Cin, Cout, k, s, p)
    gen = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(N, Cin, T, H, W, device=dev, generator=gen).bfloat16().float()
    dy = torch.randn(N, Cout, T, H, W, device=dev, generator=gen).bfloat16().float()
    w = torch.zeros(Cout, Cin, *k, device=dev, requires_grad=True)
    F.conv3d(x, w, None, s, p).backward(dy)
    x_nd, dy_nd = K.to_ndhwc(x), K.to_ndhwc(dy)
    plain = K.unpack_conv_wgrad(K.conv3d_wgrad_packed(x_nd, dy_nd, g), g)
    monkeypatch.setenv("DV_WGRAD_CLUSTER", "1")
    clustered = K.unpack_conv_wgrad(K.conv3d_wgrad_packed(x_nd, dy_nd, g), g)
    assert _rel(plain, w.grad) < 1e-4
    assert _rel(clustered, w.grad) < 1e-4
    assert _rel(clustered, plain) < 1e-5


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_trio_matches_torch(case):
    import kernel_handles as K
    torch.backends.cudnn.allow_tf32 = False
    name, N, T, H, W, Cin, Cout, k, s, p = case
    dev = "cuda:0"
    g = K.make_geom(N, T, H, W, Cin, Cout, k, s, p)
    gen = torch.Generator(device=dev).manual_seed(abs(hash(name)) % (2 ** 31))
    x = torch.randn(N, Cin, T, H, W, device=dev, generator=gen)
    w = torch.randn(Cout, Cin, *k, device=dev, generator=gen) / (Cin * k[0] * k[1] * k[2]) ** 0.5
    bias = torch.randn(Cout, device=dev, generator=gen)
    xr = x.bfloat16().float().requires_grad_(True)
    wr = w.bfloat16().float().requires_grad_(True)
    yr = F.conv3d(xr, wr, bias, s, p)
    dy = torch.randn(yr.shape, device=dev, generator=gen).bfloat16().float()
    yr.backward(dy)

    x_nd = K.to_ndhwc(x)
    assert torch.equal(K.from_ndhwc(x_nd, Cin), x.bfloat16().float())
    wf, wt = K.pack_conv_weight(w, g)
    stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
    bias_p = torch.zeros(g.Cout_p, device=dev)
    bias_p[:Cout] = bias
    y_nd = K.conv3d_fprop(x_nd, wf, g, bn_stats=stats, bias=bias_p)
    y = K.from_ndhwc(y_nd, Cout)
    assert _rel(y, yr.detach()) < 1e-2
    if g.Cout_p > Cout:
        assert bool((y_nd[..., Cout:] == 0).all())
    ys = y_nd.float()[..., :Cout].reshape(-1, Cout).double()
    torch.testing.assert_close(stats[:Cout], ys.sum(0), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(stats[g.Cout_p:g.Cout_p + Cout], (ys * ys).sum(0), rtol=1e-5, atol=1e-3)

    dy_nd = K.to_ndhwc(dy)
    dx = K.from_ndhwc(K.conv3d_dgrad(dy_nd, wt, g), Cin)
    assert _rel(dx, xr.grad) < 1e-2
    dw = K.unpack_conv_wgrad(K.conv3d_wgrad_packed(x_nd, dy_nd, g), g)
    assert _rel(dw, wr.grad) < 1e-4          # fp32 accumulate + fp32 output


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("case", [c for c in CASES if c[6] > 4 and "stem" not in c[0]], ids=lambda c: c[0])
def test_dgrad_fused_bn_reduce(case, relu):
    """dv_conv3d_dgrad_bnred_bf16: dx bit-identical to the plain dgrad, and the epilogue's sum(g), sum(g*y)
    equal to dv_bn_bwd_reduce run on that dx (and to a torch fp64 evaluation of the same sums)."""
    import ctypes
    from dualvar_b200 import _lib
    import kernel_handles as K
    name, N, T, H, W, Cin, Cout, k, s, p = case
    dev = "cuda:0"
    g = K.make_geom(N, T, H, W, Cin, Cout, k, s, p)
    gen = torch.Generator(device=dev).manual_seed(abs(hash(name)) % (2 ** 31) + 1)
    w = torch.randn(Cout, Cin, *k, device=dev, generator=gen) / (Cout * k[0] * k[1] * k[2]) ** 0.5
    _, wt = K.pack_conv_weight(w, g)
    dy = torch.randn(N, g.To, g.Ho, g.Wo, g.Cout_p, device=dev, generator=gen).bfloat16()
    dy[..., Cout:] = 0
    y_prev = torch.randn(N, T, H, W, g.Cin_p, device=dev, generator=gen).bfloat16()
    y_prev[..., Cin:] = 0
    ss = torch.randn(2 * g.Cin_p, device=dev, generator=gen)
    dx_ref = K.conv3d_dgrad(dy, wt, g)
    dx = torch.full_like(dx_ref, float("nan"))
    sums = torch.zeros(2 * g.Cin_p, dtype=torch.float64, device=dev)
    _lib.call("dv_conv3d_dgrad_bnred_bf16", _lib.ptr(dy), _lib.ptr(wt), _lib.ptr(dx), ctypes.byref(g),
              _lib.ptr(y_prev), _lib.ptr(ss) if relu else None, _lib.ptr(sums), _lib.stream_ptr())
    assert torch.equal(dx, dx_ref)
    rows = N * T * H * W
    want = torch.zeros_like(sums)
    _lib.call("dv_bn_bwd_reduce", _lib.ptr(dx), None, _lib.ptr(dx), _lib.ptr(y_prev), _lib.ptr(ss) if relu else None,
              _lib.ptr(want), rows, g.Cin_p, g.Cin_p, 0, 1 if relu else 0, _lib.stream_ptr())
    yd, dd = y_prev.double().reshape(rows, -1), dx.double().reshape(rows, -1)
    if relu:
        zf = torch.addcmul(ss[g.Cin_p:].float(), y_prev.float().reshape(rows, -1), ss[:g.Cin_p].float())  # fma like the kernels
        dd = dd * (zf > 0)
    exact = torch.cat([dd.sum(0), (dd * yd).sum(0)])
    scale = exact.abs().max().item() + 1.0
    assert (sums - exact).abs().max().item() < 1e-3 * scale      # fp32 partial sums per CTA, double across CTAs
    assert (sums - want).abs().max().item() < 2e-4 * scale


def test_conv_linearity_at_full_size():
    """BASELINE-size property check (no oracle at this size): conv(a*x1 + x2) with power-of-two a is
    exactly a*conv(x1) + conv(x2) up to bf16 rounding of the stored outputs; pad channels stay zero."""
    import kernel_handles as K
    dev = "cuda:0"
    g = K.make_geom(16, 16, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    gen = torch.Generator(device=dev).manual_seed(3)
    x1 = torch.randn(16, 16, 56, 56, 64, device=dev, generator=gen).bfloat16()
    w = torch.randn(144, 64, 1, 3, 3, device=dev, generator=gen) / 24.0
    wf, _ = K.pack_conv_weight(w, g)
    y1 = K.conv3d_fprop(x1, wf, g).float()
    y2 = K.conv3d_fprop((x1.float() * 2).bfloat16(), wf, g).float()
    assert torch.equal(y2, y1 * 2)


@pytest.mark.parametrize("kt,pt,Cout,H,W", [(1, 0, 83, 36, 44), (3, 1, 64, 36, 44), (1, 0, 83, 48, 64)])
def test_space_to_depth_stem_matches_torch(kt, pt, Cout, H, W):
    """Stride-2 7x7 stem on the space-to-depth ingest (overlapping-window TMA map) vs torch conv3d."""
    import ctypes
    from dualvar_b200 import _lib, engine as E
    import kernel_handles as K
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda:0"
    N, T = 5, 4        # 48x64 frames -> 24x32 map: the two-region tiling path
    gen = torch.Generator(device=dev).manual_seed(kt)
    x = torch.randn(N, 3, T, H, W, device=dev, generator=gen)
    w = torch.randn(Cout, 3, kt, 7, 7, device=dev, generator=gen) / (3 * kt * 49) ** 0.5
    xr = x.bfloat16().float()
    wr = w.bfloat16().float().requires_grad_(True)
    yr = F.conv3d(xr, wr, None, (1, 2, 2), (pt, 3, 3))
    dy = torch.randn(yr.shape, device=dev, generator=gen).bfloat16().float()
    yr.backward(dy)
    g = K.make_geom(N, T, H, W, 3, Cout, (kt, 7, 7), (1, 2, 2), (pt, 3, 3))
    xa = E.ingest(x, s2d=True)
    assert xa.data.shape == (N, T, H // 2, W // 2 + 3, 16)
    ws = torch.empty((g.Cout_p, kt * 4, 64), dtype=torch.bfloat16, device=dev)
    _lib.call("dv_pack_stem_weight", _lib.ptr(w), _lib.ptr(ws), ctypes.byref(g), _lib.stream_ptr())
    y_nd = torch.empty((N, g.To, g.Ho, g.Wo, g.Cout_p), dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
    _lib.call("dv_conv3d_stem_fprop_bf16", _lib.ptr(xa.data), _lib.ptr(ws), _lib.ptr(y_nd), _lib.ptr(stats), None,
              ctypes.byref(g), _lib.stream_ptr())
    assert _rel(K.from_ndhwc(y_nd, Cout), yr.detach()) < 1e-2
    ys = y_nd.float()[..., :Cout].reshape(-1, Cout).double()
    torch.testing.assert_close(stats[:Cout], ys.sum(0), rtol=1e-5, atol=1e-3)
    dws = torch.empty((g.Cout_p, kt * 4, 64), dtype=torch.float32, device=dev)
    dw = torch.empty_like(w)
    dy_nd = K.to_ndhwc(dy)
    _lib.call("dv_conv3d_stem_wgrad_bf16", _lib.ptr(xa.data), _lib.ptr(dy_nd), _lib.ptr(dws), ctypes.byref(g),
              _lib.stream_ptr())
    _lib.call("dv_unpack_stem_wgrad", _lib.ptr(dws), _lib.ptr(dw), ctypes.byref(g), ctypes.c_float(0.0),
              _lib.stream_ptr())
    assert _rel(dw, wr.grad) < 1e-4


XF_CASES = [c for c in CASES if "stem" not in c[0]] + [
    ("temporal 144->64 56x56 (CTA pairs, weight-stationary)", 12, 16, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("temporal 83->64 (Cin_p = 88)", 4, 8, 28, 28, 83, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("temporal s2 460->256 14x14", 3, 8, 14, 14, 460, 256, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
]


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("case", XF_CASES, ids=lambda c: c[0])
def test_consumer_side_batchnorm_is_bit_identical_to_the_two_pass_path(case, relu):
    """dv_conv3d_fprop_bnrelu_bf16 / dv_conv3d_wgrad_bnrelu_bf16 (the BatchNorm + ReLU of the activation applied to the
    operand tile in shared memory, csrc/bn_xform.cuh) against dv_bn_apply followed by the plain kernels: the forward
    output and its fused statistics bit for bit (same MMA order on identical operand bits), the weight gradient up to the
    fp32 atomics order of its split-K reduction. Zero padding, partial tiles, strided views and padded channels are all
    in the case list; the shift is made large so that a padding row wrongly normalised (relu(shift) != 0) shows."""
    import ctypes
    import kernel_handles as K
    from dualvar_b200 import _lib
    from dualvar_b200._lib import ptr, stream_ptr
    name, N, T, H, W, Cin, Cout, k, s, p = case
    dev = "cuda:0"
    g = K.make_geom(N, T, H, W, Cin, Cout, k, s, p)
    gen = torch.Generator(device=dev).manual_seed(abs(hash(name)) % (2 ** 31))
    y_prev = K.to_ndhwc(torch.randn(N, Cin, T, H, W, device=dev, generator=gen))         # raw output of the conv below
    ss = torch.zeros(2 * g.Cin_p, device=dev)
    ss[:Cin] = torch.rand(Cin, device=dev, generator=gen) + 0.5                            # scale
    ss[g.Cin_p:g.Cin_p + Cin] = torch.randn(Cin, device=dev, generator=gen) + 0.7         # shift (mostly positive)
    w = torch.randn(Cout, Cin, *k, device=dev, generator=gen) / (Cin * k[0] * k[1] * k[2]) ** 0.5
    wf, wt = K.pack_conv_weight(w, g)
    rows = y_prev.numel() // g.Cin_p
    z = torch.empty_like(y_prev)
    _lib.call("dv_bn_apply", ptr(y_prev), ptr(ss), None, None, None, ptr(z), rows, g.Cin_p, g.Cin_p, 0, 1 if relu else 0,
              stream_ptr())
    st_a = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
    y_a = K.conv3d_fprop(z, wf, g, bn_stats=st_a)
    st_b = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
    y_b = torch.empty_like(y_a)
    _lib.call("dv_conv3d_fprop_bnrelu_bf16", ptr(y_prev), ptr(ss), 1 if relu else 0, ptr(wf), ptr(y_b), ptr(st_b), None,
              ctypes.byref(g), stream_ptr())
    assert torch.equal(y_a, y_b), (y_a.float() - y_b.float()).abs().max().item()
    torch.testing.assert_close(st_b, st_a, rtol=1e-12, atol=1e-6)
    dy = torch.randn(y_a.shape, device=dev, generator=gen).bfloat16()
    if g.Cout_p > Cout:
        dy[..., Cout:] = 0
    dw_a = K.conv3d_wgrad_packed(z, dy, g)
    dw_b = torch.empty_like(dw_a)
    _lib.call("dv_conv3d_wgrad_bnrelu_bf16", ptr(y_prev), ptr(ss), 1 if relu else 0, ptr(dy), ptr(dw_b), ctypes.byref(g),
              stream_ptr())
    assert ((dw_a - dw_b).abs().max() / dw_a.abs().max()).item() < 2e-5


@pytest.mark.parametrize("geom", [(4, 8, 28, 56, 64, 144), (3, 4, 56, 56, 64, 40), (3, 16, 42, 24, 57, 128)])
def test_kh_stacked_dgrad_matches_the_tap_by_tap_dgrad(geom):
    """dv_conv3d_dgrad_stack_bf16 (kh taps stacked along N, one MMA per kw and K step on the unshifted box, row-shifted sum in
    the epilogue) against dv_conv3d_dgrad_bf16: same products in another fp32 summation order - equal up to one bf16
    last place on a few elements; its fused BatchNorm-backward sums against the fused tap-by-tap launch."""
    import ctypes
    import kernel_handles as K
    from dualvar_b200 import _lib
    n, t, h, w, ci, co = geom
    dev = "cuda:0"
    g = K.make_geom(n, t, h, w, ci, co, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    assert _lib.load().dv_conv3d_dgrad_stack_ok(ctypes.byref(g)) == 1
    gen = torch.Generator(device=dev).manual_seed(n + h + co)
    wt_f = torch.randn(co, ci, 1, 3, 3, device=dev, generator=gen) / (ci * 9) ** 0.5
    _, wt = K.pack_conv_weight(wt_f, g)
    ws = wt.view(g.Cin_p, 3, 3, g.Cout_p).flip(1).permute(1, 0, 2, 3).reshape(3 * g.Cin_p, 3, g.Cout_p).contiguous()
    dy = torch.randn(n, t, h, w, g.Cout_p, device=dev, generator=gen).bfloat16()
    dy[..., co:] = 0
    ref = K.conv3d_dgrad(dy, wt, g)
    dx = torch.full_like(ref, float("nan"))
    _lib.call("dv_conv3d_dgrad_stack_bf16", _lib.ptr(dy), _lib.ptr(ws), _lib.ptr(dx), ctypes.byref(g), None, None, None,
              _lib.stream_ptr())
    diff = (dx.float() - ref.float()).abs()
    assert not torch.isnan(dx.float()).any()
    assert diff.max().item() <= 2.0 ** -7 * ref.float().abs().max().item()
    assert (diff > 0).float().mean().item() < 5e-3
    if g.Cin_p > ci:
        assert bool((dx[..., ci:] == 0).all())
    y_prev = torch.randn(n, t, h, w, g.Cin_p, device=dev, generator=gen).bfloat16()
    y_prev[..., ci:] = 0
    ss = torch.randn(2 * g.Cin_p, device=dev, generator=gen)
    s_ref = torch.zeros(2 * g.Cin_p, dtype=torch.float64, device=dev)
    s_new = torch.zeros_like(s_ref)
    dx1, dx2 = torch.empty_like(ref), torch.empty_like(ref)
    _lib.call("dv_conv3d_dgrad_bnred_bf16", _lib.ptr(dy), _lib.ptr(wt), _lib.ptr(dx1), ctypes.byref(g), _lib.ptr(y_prev),
              _lib.ptr(ss), _lib.ptr(s_ref), _lib.stream_ptr())
    _lib.call("dv_conv3d_dgrad_stack_bf16", _lib.ptr(dy), _lib.ptr(ws), _lib.ptr(dx2), ctypes.byref(g), _lib.ptr(y_prev),
              _lib.ptr(ss), _lib.ptr(s_new), _lib.stream_ptr())
    assert torch.equal(dx2, dx)
    assert (s_new - s_ref).abs().max().item() <= 2e-3 * (s_ref.abs().max().item() + 1.0)


def test_kh_stacked_dgrad_refuses_other_geometries():
    import ctypes
    import kernel_handles as K
    from dualvar_b200 import _lib
    for geom, k, s, p in [((4, 8, 28, 56, 128, 144), (1, 3, 3), (1, 1, 1), (0, 1, 1)),      # 128 gradient columns
                          ((4, 8, 28, 56, 64, 144), (3, 1, 1), (1, 1, 1), (1, 0, 0)),       # temporal filter
                          ((4, 8, 56, 56, 64, 144), (1, 3, 3), (1, 2, 2), (0, 1, 1)),       # strided
                          ((4, 8, 30, 56, 64, 144), (1, 3, 3), (1, 1, 1), (0, 1, 1)),       # H not a multiple of 14
                          ((1, 1, 28, 56, 64, 144), (1, 3, 3), (1, 1, 1), (0, 1, 1))]:      # too few tiles
        g = K.make_geom(*geom, k, s, p)
        assert _lib.load().dv_conv3d_dgrad_stack_ok(ctypes.byref(g)) == 0
