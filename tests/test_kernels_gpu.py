"""GPU: memory-bound kernels and fp32 objectives through the C ABI vs torch / the oracle.
fp32 pieces (heads, normalise, losses) are held to the north-star's fp32 tolerance (1e-4 relative);
bf16 activations to 1e-2."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda:0"


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def _rel2(a, b):
    """L2-relative error: the right yardstick through ReLU, whose mask flips for the handful of
    pre-activations that bf16 rounding moves across zero (each flip changes that element wholesale)."""
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_objectives_match_golden_reference_vectors(golden_dir):
    """The CUDA objectives against the fixtures generated from the REAL reference."""
    from dualvar_b200 import objectives as O
    g = np.load(os.path.join(golden_dir, "objectives.npz"))
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    f = t("ntx_in").requires_grad_(True)
    ret, hits = O.nt_xent(f, 0.07, False)
    ret["clip_contrast_loss"].backward()
    torch.testing.assert_close(ret["clip_logits"], t("ntx_logits"), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ret["clip_contrast_loss"], t("ntx_loss"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(f.grad, t("ntx_grad"), rtol=1e-3, atol=1e-5)
    s = t("tc_in").requires_grad_(True)
    ret, _ = O.tc_loss(s, 0.07, False)
    ret["tc_contrast_loss"].backward()
    torch.testing.assert_close(ret["tc_logits"], t("tc_logits"), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ret["tc_contrast_loss"], t("tc_loss"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(s.grad, t("tc_grad"), rtol=1e-3, atol=1e-5)
    p = t("rank_in")
    a, b = p[:, :, 0].contiguous().requires_grad_(True), p[:, :, 1].contiguous().requires_grad_(True)
    ret, _ = O.rank_loss(a, b, 0.05, 0.5, 5.0, "x_")
    ret["x_margin_contrast_loss"].backward()
    torch.testing.assert_close(ret["x_margin_logits"], t("rank_logits"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ret["x_margin_contrast_loss"], t("rank_loss"), rtol=1e-5, atol=1e-6)
    gref = t("rank_grad")
    torch.testing.assert_close(a.grad, gref[:, :, 0], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(b.grad, gref[:, :, 1], rtol=1e-3, atol=1e-5)
    p3 = t("rank3_in")
    ret, _ = O.rank_loss(p3[:, :, 0].contiguous(), p3[:, :, 1].contiguous(), 0.05, 0.5, 5.0, "x_")
    torch.testing.assert_close(ret["x_margin_logits"], t("rank3_logits"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ret["x_margin_contrast_loss"], t("rank3_loss"), rtol=1e-5, atol=1e-6)
    ret, _ = O.rank_loss(p[:, :, 0].contiguous(), p[:, :, 1].contiguous(), 0.05, 0.5, None, "x_")
    torch.testing.assert_close(ret["x_margin_contrast_loss"], t("mrank_loss"), rtol=1e-5, atol=1e-6)
    q = t("moco_q").requires_grad_(True)
    ret, _ = O.queue_contrast(q, t("moco_k"), t("moco_queue"), 0.07, "clip_")
    ret["clip_contrast_loss"].backward()
    torch.testing.assert_close(ret["clip_logits"], t("moco_logits"), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ret["clip_contrast_loss"], t("moco_loss"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(q.grad, t("moco_grad"), rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("n,d", [(64, 128), (256, 128), (37, 64)])
def test_nt_xent_and_tc_match_oracle_at_batch_sizes(n, d):
    from dualvar_b200 import objectives as O
    from oracle import objectives as OO
    gen = torch.Generator(device=dev).manual_seed(n)
    f = F.normalize(torch.randn(n, 2, d, device=dev, generator=gen), dim=-1)
    f1, f2 = f.clone().requires_grad_(True), f.clone().requires_grad_(True)
    ret, hits = O.nt_xent(f1, 0.07, False)
    logits, labels, loss = OO.nt_xent(f2, 0.07)
    ret["clip_contrast_loss"].backward(); loss.backward()
    assert ret["clip_logits"].shape == logits.shape == (2 * n, 2 * n - 1)
    torch.testing.assert_close(ret["clip_logits"], logits, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ret["clip_contrast_loss"], loss, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(f1.grad, f2.grad, rtol=1e-3, atol=1e-6)
    top1, top5 = OO.topk_accuracy(logits, labels, (1, 5))
    assert abs(hits[0].item() / (2 * n) - top1.item()) < 1e-6 and abs(hits[1].item() / (2 * n) - top5.item()) < 1e-6
    s = F.normalize(torch.randn(n, 2, 2, 64, device=dev, generator=gen), dim=-1)
    s1, s2 = s.clone().requires_grad_(True), s.clone().requires_grad_(True)
    ret, _ = O.tc_loss(s1, 0.07, False)
    logits, _, loss = OO.tc_loss(s2, 0.07)
    ret["tc_contrast_loss"].backward(); loss.backward()
    torch.testing.assert_close(ret["tc_logits"], logits, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ret["tc_contrast_loss"], loss, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(s1.grad, s2.grad, rtol=1e-3, atol=1e-6)


def test_heads_and_normalize_match_torch():
    from dualvar_b200 import objectives as O
    gen = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(192, 512, device=dev, generator=gen)
    c1, c2 = nn.Conv3d(512, 512, 1).to(dev), nn.Conv3d(512, 128, 1).to(dev)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y = O.l2norm(O.linear(O.linear(x1, c1, relu=True), c2))
    yr = F.normalize(c2(F.relu(c1(x2.view(192, 512, 1, 1, 1)))), dim=1).view(192, 128)
    torch.testing.assert_close(y, yr, rtol=1e-4, atol=1e-5)
    gy = torch.randn_like(y)
    gw = torch.autograd.grad(y, [x1, c1.weight, c1.bias, c2.weight, c2.bias], gy)
    gr = torch.autograd.grad(yr, [x2, c1.weight, c1.bias, c2.weight, c2.bias], gy)
    for a, b in zip(gw, gr):
        assert _rel(a, b) < 1e-4


def test_segment_permutation_roundtrip_and_reference_semantics():
    from dualvar_b200 import objectives as O
    from oracle import objectives as OO
    gen = torch.Generator(device=dev).manual_seed(2)
    x = torch.randn(16, 4, 64, device=dev, generator=gen)
    perms = np.array([np.random.RandomState(i).permutation(4) for i in range(16)])
    p = torch.from_numpy(perms.astype(np.int32)).to(dev)
    y = O.PermuteSegmentsFn.apply(x, p)
    assert torch.equal(y, OO.calibrate_segments(x, perms))
    # idempotence-style property: scatter then gather with the same permutation is the identity
    from dualvar_b200 import _lib
    back = torch.empty_like(x)
    _lib.call("dv_permute_segments", _lib.ptr(y), _lib.ptr(back), _lib.ptr(p), 16, 4, 64, 1, _lib.stream_ptr())
    assert torch.equal(back, x)


def test_ingest_matches_reference_transform_and_shuffle():
    """RawClips ingest == Normalize + view + transpose (pretrain.py:386-389) and the segment gather of
    model/simclr.py:378-383, bit-exact after bf16 rounding."""
    from dualvar_b200 import engine as E
    from oracle import objectives as OO
    gen = torch.Generator(device=dev).manual_seed(3)
    B, T, H, W = 4, 8, 12, 10
    frames = torch.rand(B, 3, 3 * T, H, W, device=dev, generator=gen)
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1, 1)
    block = ((frames - mean) / std).view(B, 3, 3, T, H, W).transpose(1, 2).contiguous()     # (B, V, C, T, H, W)
    want = block.reshape(B * 3, 3, T, H, W).permute(0, 2, 3, 4, 1)
    got = E.ingest(E.RawClips(frames, 3)).data
    assert got.shape == (B * 3, T, H, W, 8) and bool((got[..., 3:] == 0).all())
    torch.testing.assert_close(got[..., :3].float(), want.bfloat16().float(), rtol=8e-3, atol=1e-6)
    got_b = E.ingest(block).data
    assert torch.equal(got_b[..., :3].float(), want.bfloat16().float())
    perms = np.array([np.random.RandomState(i).permutation(2) for i in range(B)])
    shuf = OO.shuffle_segments(block[:, 2].contiguous(), perms).permute(0, 2, 3, 4, 1)
    got_s = E.ingest(block, first_view=2, n_views=1, perm=torch.from_numpy(perms.astype(np.int32)).to(dev),
                     n_series=2).data
    assert torch.equal(got_s[..., :3].float(), shuf.bfloat16().float())


@pytest.mark.parametrize("s2d", [False, True])
def test_uint8_ingest_equals_totensor_then_float_ingest(s2d):
    """uint8 frames (decoded images): dv_ingest_clips_u8 == ToTensor (x / 255 in fp32, utils/augmentation.py:361-364)
    followed by the fp32 ingest, bit for bit, in both output layouts and with the segment shuffle."""
    from dualvar_b200 import engine as E
    gen = torch.Generator(device=dev).manual_seed(9)
    B, T, H, W = 3, 8, 12, 10
    u8 = torch.randint(0, 256, (B, 3, 3 * T, H, W), device=dev, generator=gen, dtype=torch.uint8)
    as_float = u8.float().div(255)
    perm = torch.from_numpy(np.array([np.random.RandomState(i).permutation(2) for i in range(B)], dtype=np.int32)).to(dev)
    for kw in (dict(), dict(first_view=2, n_views=1, perm=perm, n_series=2)):
        a = E.ingest(E.RawClips(u8, 3), s2d=s2d, **kw).data
        b = E.ingest(E.RawClips(as_float, 3), s2d=s2d, **kw).data
        assert a.dtype == torch.bfloat16 and torch.equal(a, b)


def test_conv_bn_relu_residual_block_forward_backward():
    """One fused unit (conv -> BN(train) -> (+res) -> ReLU) forward and backward vs torch fp32 on the
    same bf16-rounded input; checks running statistics too."""
    from dualvar_b200 import engine as E
    torch.manual_seed(4)
    gen = torch.Generator(device=dev).manual_seed(4)
    N, C, T, H, W = 6, 64, 4, 16, 16
    conv = nn.Conv3d(C, C, (1, 3, 3), padding=(0, 1, 1), bias=False).to(dev)
    bn = nn.BatchNorm3d(C).to(dev)
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.2)
    conv_r, bn_r = nn.Conv3d(C, C, (1, 3, 3), padding=(0, 1, 1), bias=False).to(dev), nn.BatchNorm3d(C).to(dev)
    conv_r.load_state_dict(conv.state_dict()); bn_r.load_state_dict(bn.state_dict())
    conv_r.weight.data = conv_r.weight.data.bfloat16().float()
    x = torch.randn(N, C, T, H, W, device=dev, generator=gen).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    yr = F.relu(bn_r(conv_r(xr)) + xr)
    gy = torch.randn_like(yr).bfloat16().float()
    yr.backward(gy)

    ctx = E.Context(training=True)
    import kernel_handles as K
    xa = E.Act(K.to_ndhwc(x), C)
    out = E.activate(ctx, E.conv_stats(ctx, xa, conv, bn), res=xa)
    E.flush_batch_counters(ctx)           # what the end of a backbone pass does (one multi-tensor increment)
    y = K.from_ndhwc(out.data, C)
    assert _rel(y, yr.detach()) < 1e-2
    out.grad = K.to_ndhwc(gy)
    E.run_backward(ctx)
    assert xa.grad2 is not None          # main-path and shortcut gradients are kept as a pair ...
    dx = K.from_ndhwc(E._materialize_grad(xa), C)   # ... and summed on demand
    assert _rel2(dx, xr.grad) < 4e-2      # ~sqrt(mask-flip fraction 2e-4) + bf16 rounding
    assert ((dx - xr.grad).abs() > 0.05 * xr.grad.abs().max()).float().mean().item() < 1e-3
    assert _rel2(ctx.param_grads[id(conv.weight)], conv_r.weight.grad) < 4e-2
    assert _rel2(ctx.param_grads[id(bn.weight)], bn_r.weight.grad) < 4e-2
    assert _rel2(ctx.param_grads[id(bn.bias)], bn_r.bias.grad) < 4e-2
    assert _rel(bn.running_mean, bn_r.running_mean) < 1e-2 and _rel(bn.running_var, bn_r.running_var) < 1e-2
    assert int(bn.num_batches_tracked) == 1


def test_maxpool_forward_backward_matches_torch():
    from dualvar_b200 import engine as E
    import kernel_handles as K
    gen = torch.Generator(device=dev).manual_seed(5)
    for kernel, stride, pad in [((1, 2, 2), (1, 2, 2), (0, 0, 0)), ((2, 2, 2), (2, 2, 2), (0, 0, 0)),
                                ((3, 3, 3), (1, 1, 1), (1, 1, 1)), ((1, 3, 3), (1, 2, 2), (0, 1, 1)),
                                ((3, 3, 3), (2, 2, 2), (1, 1, 1))]:
        x = torch.randn(2, 24, 6, 13, 14, device=dev, generator=gen).bfloat16().float()
        xr = x.clone().requires_grad_(True)
        yr = F.max_pool3d(xr, kernel, stride, pad)
        gy = torch.randn_like(yr).bfloat16().float()
        yr.backward(gy)
        ctx = E.Context(training=True)
        xa = E.Act(K.to_ndhwc(x), 24)
        out = E.max_pool(ctx, xa, kernel, stride, pad)
        assert torch.equal(K.from_ndhwc(out.data, 24), yr.detach())
        out.grad = K.to_ndhwc(gy)
        E.run_backward(ctx)
        assert _rel2(K.from_ndhwc(xa.grad, 24), xr.grad) < 1e-2


def test_maxpool_backward_tie_rule_matches_torch():
    """ReLU'd inputs are full of exact ties (zeros): the gradient must go to the FIRST maximum of each window in
    (t,h,w) scan order, as ATen's max_pool3d does - for the recorded-argmax path and the re-scanning C-ABI entry."""
    import ctypes
    from dualvar_b200 import _lib, engine as E
    import kernel_handles as K
    gen = torch.Generator(device=dev).manual_seed(6)
    for kernel, stride, pad in [((3, 3, 3), (1, 1, 1), (1, 1, 1)), ((1, 3, 3), (1, 2, 2), (0, 1, 1))]:
        x = torch.relu(torch.randn(2, 16, 5, 12, 11, device=dev, generator=gen) - 0.8).bfloat16().float()
        xr = x.clone().requires_grad_(True)
        yr = F.max_pool3d(xr, kernel, stride, pad)
        gy = (torch.randint(-3, 4, yr.shape, device=dev, generator=gen).float() / 4)    # exactly representable sums
        yr.backward(gy)
        ctx = E.Context(training=True)
        xa = E.Act(K.to_ndhwc(x), 16)
        out = E.max_pool(ctx, xa, kernel, stride, pad)
        assert torch.equal(K.from_ndhwc(out.data, 16), yr.detach())
        out.grad = K.to_ndhwc(gy)
        E.run_backward(ctx)
        assert torch.equal(K.from_ndhwc(xa.grad, 16), xr.grad)
        # the index-free entry point (re-scans the window for earlier ties) gives the same answer
        N, T, H, W, Cp = xa.shape5
        To, Ho, Wo = out.data.shape[1:4]
        geom = (ctypes.c_int32 * 17)(N, T, H, W, To, Ho, Wo, Cp, *kernel, *stride, *pad)
        dx = torch.empty_like(xa.data)
        _lib.call("dv_maxpool3d_bwd", _lib.ptr(xa.data), _lib.ptr(out.data), _lib.ptr(K.to_ndhwc(gy)), _lib.ptr(dx), geom,
                  _lib.stream_ptr())
        assert torch.equal(K.from_ndhwc(dx, 16), xr.grad)


def test_fused_sgd_matches_torch_sgd():
    """dualvar_b200.optim.SGD (one launch for all tensors) vs torch.optim.SGD, three steps with an lr change."""
    from dualvar_b200.optim import SGD
    torch.manual_seed(0)
    shapes = [(83, 3, 1, 7, 7), (64,), (144, 64, 1, 3, 3), (10001,), (7,)]
    ps = [torch.randn(s, device=dev).requires_grad_(True) for s in shapes]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    ours = SGD([{"params": p} for p in ps], lr=0.003, momentum=0.9, weight_decay=1e-4)     # pretrain.py:262-272
    ref = torch.optim.SGD([{"params": q} for q in qs], lr=0.003, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        if step == 2:
            for grp in ours.param_groups + ref.param_groups:
                grp["lr"] = 0.0003
        ours.step(); ref.step()
        for p, q in zip(ps, qs):
            # fp32 both sides; the kernel contracts g + wd*p and mu*buf + g into FMAs (torch's foreach ops round
            # each product), so results agree to an ulp or two of the operands, not bit for bit
            torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-6)
    for p, q in zip(ps, qs):
        torch.testing.assert_close(ours.state[p]["momentum_buffer"], ref.state[q]["momentum_buffer"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("N,S,C,ld,coff", [(5, 300, 64, 256, 64), (3, 77, 208, 512, 96), (48, 64, 24, 24, 0), (2, 4096, 384, 1024, 256)])
def test_self_gating_kernels_match_torch(N, S, C, ld, coff):
    """S3D-G SelfGating on a channel slice of the concat tensor (backbone/s3dg.py:68-78): slice mean, Linear + sigmoid,
    in-place scale, and the backward of all three, kernel by kernel against torch."""
    import ctypes
    from dualvar_b200 import _lib
    from dualvar_b200._lib import ptr, stream_ptr
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(N * 1000 + C)
    cat = torch.randn(N, S, ld, device=dev, generator=g).bfloat16()
    W = torch.randn(C, C, device=dev, generator=g) / C ** 0.5
    b = torch.randn(C, device=dev, generator=g)
    # forward
    mean = torch.empty(N, C, device=dev)
    _lib.call("dv_slice_mean", ptr(cat), ptr(mean), N, S, C, ld, coff, stream_ptr())
    sl = cat[:, :, coff:coff + C].float()
    torch.testing.assert_close(mean, sl.mean(1), rtol=1e-5, atol=1e-5)
    w = torch.empty(N, C, device=dev)
    _lib.call("dv_gate_fc_fwd", ptr(mean), ptr(W), ptr(b), ptr(w), N, C, stream_ptr())
    w_ref = torch.sigmoid(mean @ W.t() + b)
    torch.testing.assert_close(w, w_ref, rtol=1e-5, atol=1e-6)
    scaled = cat.clone()
    _lib.call("dv_gate_scale", ptr(scaled), ptr(w), N, S, C, ld, coff, stream_ptr())
    assert torch.equal(scaled[:, :, coff:coff + C], (sl * w[:, None, :]).bfloat16())
    keep = torch.ones(ld, dtype=torch.bool, device=dev); keep[coff:coff + C] = False
    assert torch.equal(scaled[:, :, keep], cat[:, :, keep])                      # other slices untouched
    # backward: z = relu(scale*y + shift) recomputed from the branch's raw output y
    y = torch.randn(N, S, C, device=dev, generator=g).bfloat16()
    ss = torch.cat([torch.rand(C, device=dev, generator=g) + 0.5, torch.randn(C, device=dev, generator=g)])
    dout = torch.randn(N, S, ld, device=dev, generator=g).bfloat16()
    z = torch.relu(y.float() * ss[:C] + ss[C:])
    dw = torch.empty(N, C, device=dev)
    _lib.call("dv_gate_bwd_reduce", ptr(dout), ptr(y), ptr(ss), ptr(dw), N, S, C, C, ld, coff, stream_ptr())
    dsl = dout[:, :, coff:coff + C].float()
    torch.testing.assert_close(dw, (dsl * z).sum(1), rtol=2e-4, atol=2e-3)
    mean_r = mean.clone().requires_grad_(True); W_r = W.clone().requires_grad_(True); b_r = b.clone().requires_grad_(True)
    torch.sigmoid(mean_r @ W_r.t() + b_r).backward(dw)
    dpre, gW, gb, dmean = torch.empty(N, C, device=dev), torch.empty_like(W), torch.empty_like(b), torch.empty(N, C, device=dev)
    _lib.call("dv_gate_fc_bwd", ptr(dw), ptr(w), ptr(mean), ptr(W), ptr(dpre), ptr(gW), ptr(gb), ptr(dmean), N, C, stream_ptr())
    scale = dw.abs().max().item()
    torch.testing.assert_close(gW, W_r.grad, rtol=1e-4, atol=1e-4 * scale)
    torch.testing.assert_close(gb, b_r.grad, rtol=1e-4, atol=1e-4 * scale)
    torch.testing.assert_close(dmean, mean_r.grad, rtol=1e-4, atol=1e-4 * scale)
    dz = torch.empty(N, S, C, device=dev, dtype=torch.bfloat16)
    _lib.call("dv_gate_bwd_apply", ptr(dout), ptr(w), ptr(dmean), ptr(dz), N, S, C, C, ld, coff, stream_ptr())
    want = (w[:, None, :] * dsl + dmean[:, None, :] / S)
    assert ((dz.float() - want).abs().max() / want.abs().max()).item() < 1e-2      # bf16 output


@pytest.mark.parametrize("M,N,K", [(192, 512, 512), (37, 129, 70), (1, 1, 1), (256, 128, 33), (300, 600, 96), (64, 16385, 128)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_sgemm_ragged_sizes_and_transposes_match_torch(M, N, K, ta, tb):
    """dv_sgemm through the C ABI (small double-buffered kernel below 2 x 148 tiles of 64 x 64, the 64 x 64 kernel above):
    every transpose combination, ragged M / N / K, bias + ReLU epilogue and beta accumulation, against torch fp32."""
    import ctypes
    from dualvar_b200 import _lib
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    A = torch.randn((K, M) if ta else (M, K), device=dev, generator=gen)
    B = torch.randn((N, K) if tb else (K, N), device=dev, generator=gen)
    bias = torch.randn(N, device=dev, generator=gen)
    C0 = torch.randn(M, N, device=dev, generator=gen)
    C = C0.clone()
    _lib.call("dv_sgemm", ta, tb, M, N, K, ctypes.c_float(0.5), _lib.ptr(A), A.shape[1], _lib.ptr(B), B.shape[1],
              ctypes.c_float(0.25), _lib.ptr(C), N, _lib.ptr(bias), 1, _lib.stream_ptr())
    want = F.relu(0.5 * ((A.t() if ta else A).double() @ (B.t() if tb else B).double()) + bias.double() + 0.25 * C0.double())
    assert (C.double() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item()) * max(1, K) ** 0.5


@pytest.mark.parametrize("R,C,d,col0,dmajor", [(128, 128, 128, 0, 0), (8, 24, 64, 0, 0), (64, 16384, 128, 1, 1),
                                               (37, 301, 40, 1, 1), (5, 130, 16, 2, 0)])
def test_fused_similarity_logsumexp_matches_a_torch_evaluation(R, C, d, col0, dmajor):
    """dv_sim_ce_fwd + dv_sim_ce_finish (similarity GEMM fused with the row-wise log-sum-exp / cross-entropy,
    model/simclr.py:198-221, model/moco.py:426-438): S, logits in the reference's column order, loss, top-1 / top-5
    counts and dLoss/dS against a float64 torch evaluation - NT-Xent layout (self column dropped, row-major columns),
    MoCo layout (leading positive column, d-major queue) and ragged sizes."""
    import ctypes
    from dualvar_b200 import _lib
    from dualvar_b200.objectives import _sim_blocks
    gen = torch.Generator(device=dev).manual_seed(R + C + d)
    a = F.normalize(torch.randn(R, d, device=dev, generator=gen), dim=1)
    bmat = F.normalize(torch.randn(C, d, device=dev, generator=gen), dim=1)           # column features [C][d]
    b_arg = bmat.t().contiguous() if dmajor else bmat
    ldb = C if dmajor else d
    Ct = col0 + C
    inv_T = 1.0 / 0.07
    S = torch.zeros(R, Ct, device=dev)
    lead = torch.randn(R, col0, device=dev, generator=gen) if col0 else None
    if col0:
        S[:, :col0] = lead
    use_self = col0 == 0
    self_col = (torch.arange(R, device=dev) % Ct).to(torch.int32) if use_self else None
    pos_col = ((torch.arange(R, device=dev) * 7 + 3) % Ct).to(torch.int32)
    if use_self:
        pos_col = torch.where(pos_col == self_col, (pos_col + 1) % Ct, pos_col).to(torch.int32)
    n_out = Ct - (1 if use_self else 0)
    logits = torch.full((R, n_out), float("nan"), device=dev)
    partials = torch.empty(R, _sim_blocks(C), 2, device=dev)
    loss_sum = torch.zeros(1, device=dev)
    hits = torch.zeros(2, dtype=torch.int32, device=dev)
    _lib.call("dv_sim_ce_fwd", _lib.ptr(a), d, _lib.ptr(b_arg), ldb, dmajor, R, C, d, _lib.ptr(S), Ct, col0, _lib.ptr(logits),
              n_out, _lib.ptr(self_col), _lib.ptr(pos_col), ctypes.c_float(inv_T), _lib.ptr(partials), _lib.stream_ptr())
    S_raw = S.clone()
    _lib.call("dv_sim_ce_finish", _lib.ptr(S), Ct, R, C, col0, _lib.ptr(partials), _lib.ptr(logits), n_out,
              _lib.ptr(self_col), _lib.ptr(pos_col), ctypes.c_float(inv_T), ctypes.c_float(1.0 / R), _lib.ptr(loss_sum),
              _lib.ptr(hits), _lib.stream_ptr())
    # reference evaluation in float64
    full = torch.cat([lead.double(), a.double() @ bmat.double().t()], dim=1) if col0 else a.double() @ bmat.double().t()
    torch.testing.assert_close(S_raw.double(), full, rtol=0, atol=2e-6)
    z = S_raw.double() * inv_T                     # the kernel's own products: isolates the softmax part
    loss, top1, top5 = 0.0, 0, 0
    dS = torch.zeros_like(z)
    for r in range(R):
        keep = torch.ones(Ct, dtype=torch.bool, device=dev)
        if use_self:
            keep[int(self_col[r])] = False
        p = int(pos_col[r])
        zr = z[r][keep]
        lse = torch.logsumexp(zr, 0)
        loss += float(lse - z[r, p])
        above = int(((z[r] > z[r, p]) & keep).sum())
        top1 += above < 1
        top5 += above < 5
        sm = torch.zeros(Ct, dtype=torch.float64, device=dev)
        sm[keep] = torch.softmax(zr, 0)
        sm[p] -= 1.0
        dS[r] = sm * inv_T / R
        order = [p] + [c for c in range(Ct) if c != p and keep[c]]
        torch.testing.assert_close(logits[r].double(), z[r][order], rtol=1e-6, atol=1e-5)
    assert abs(float(loss_sum) - loss) <= 2e-5 * max(1.0, abs(loss))
    assert hits.tolist() == [top1, top5]
    torch.testing.assert_close(S.double(), dS, rtol=1e-4, atol=1e-7)
