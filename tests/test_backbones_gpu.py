"""GPU: C3D / S3D / S3D-G drop-in backbones vs the oracle. Deep BN stacks at test-sized extents amplify
bf16 rounding (the final S3D map is 2x2x2), so the yardstick is the oracle itself under torch's bf16
autocast: the product's error must track it stage by stage (a structural bug — wrong concat offset,
gating, pooling — shows up as a jump at the first affected stage)."""
import copy
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda:0"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _pair(name):
    from dualvar_b200 import backbones as PB
    from oracle import backbones as OB
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0); np.random.seed(0); random.seed(0)
    ref, _ = OB.select_backbone(name)
    ref = ref.to(dev).train()
    prod, _ = PB.select_backbone(name)
    prod.load_state_dict(ref.state_dict())
    return ref, prod.to(dev).train()


STAGES = ["Conv_1a", "MaxPool_2a", "Conv_2b", "Conv_2c", "MaxPool_3a", "Mixed_3b", "Mixed_3c", "MaxPool_4a",
          "Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f", "MaxPool_5a", "Mixed_5b", "Mixed_5c"]


@pytest.mark.parametrize("name", ["s3d", "s3dg"])
def test_s3d_stagewise_error_tracks_bf16_autocast(name):
    from dualvar_b200 import engine as E, s3dg as PS
    ref, prod = _pair(name)
    x = torch.randn(8, 3, 16, 64, 64, device=dev)
    ref_out, ref_ac, prod_out = {}, {}, {}

    def grab(store, cast):
        hs = []
        for n in STAGES:
            def h(mod, inp, out, n=n):
                store.setdefault(n, out.detach().float() if cast else out.detach())
            hs.append(getattr(ref, n).register_forward_hook(h))
        return hs

    with torch.no_grad():
        hs = grab(ref_out, False); ref(x); [h.remove() for h in hs]
        hs = grab(ref_ac, True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            copy.deepcopy(ref)
            ref(x)
        [h.remove() for h in hs]
    # undo the statistics updates of the two oracle passes before running the product
    for n in STAGES:
        m = getattr(prod, n)
        if hasattr(m, "run"):
            def wrap(orig, n=n):
                def run(ctx, xa, *a, **k):
                    r = orig(ctx, xa, *a, **k)
                    prod_out[n] = E.to_ncdhw(r[0] if isinstance(r, tuple) else r)
                    return r
                return run
            m.run = wrap(m.run)
    orig_pool = PS.S3D._pool

    def pool(ctx, xa, mod):
        o = orig_pool(ctx, xa, mod)
        for n in STAGES:
            if getattr(prod, n) is mod:
                prod_out[n] = E.to_ncdhw(o)
        return o
    PS.S3D._pool = staticmethod(pool)
    try:
        with torch.no_grad():
            y = prod(x)
    finally:
        PS.S3D._pool = staticmethod(orig_pool)
    assert y.shape == ref_out["Mixed_5c"].shape == (8, 1024, 2, 2, 2)
    for n in STAGES:
        ours, yard = _rel(prod_out[n], ref_out[n]), _rel(ref_ac[n], ref_out[n])
        assert ours <= 1.35 * yard + 5e-3, (n, ours, yard)
    assert _rel(prod_out["Conv_1a"], ref_out["Conv_1a"]) < 1e-2      # stem: S2D path + (7,1,1) stride-2 conv


@pytest.mark.parametrize("name,shape", [("c3d", (8, 3, 8, 64, 64)), ("s3d", (4, 3, 16, 64, 64)),
                                        ("s3dg", (4, 3, 16, 64, 64)), ("r2d3d18", (6, 3, 4, 96, 96))])
def test_backbone_forward_backward_vs_autocast_yardstick(name, shape):
    ref, prod = _pair(name)
    x = torch.randn(*shape, device=dev)
    yr, yp = ref(x), prod(x)
    assert yp.shape == yr.shape and torch.isfinite(yp).all()
    g = torch.randn_like(yr)
    yr.backward(g); yp.backward(g)
    ref2 = copy.deepcopy(ref)
    for p in ref2.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = ref2(x)
    ya.float().backward(g)
    # (both errors are O(0.7) for S3D at this size and fluctuate run to run with the atomics order)
    assert _rel(yp, yr) <= 1.5 * _rel(ya, yr) + 2e-2
    ours, yard = [], []
    for (n, pr), (_, pa), (_, pp) in zip(ref.named_parameters(), ref2.named_parameters(), prod.named_parameters()):
        assert pp.grad is not None and torch.isfinite(pp.grad).all(), n
        if pr.grad.norm() < 1e-6:           # conv bias in front of train-mode BN: exactly zero gradient
            continue
        ours.append(_rel(pp.grad, pr.grad)); yard.append(_rel(pa.grad, pr.grad))
    ours.sort(); yard.sort()
    assert ours[len(ours) // 2] <= 1.5 * yard[len(yard) // 2] + 0.05
    for (n, br), (_, bp) in zip(ref.named_buffers(), prod.named_buffers()):
        if not br.dtype.is_floating_point:
            assert int(bp) == int(br), n


def test_r2d3d18_eval_matches_golden_and_full_size_shape():
    """r2d3d18 in eval mode (folded running statistics) on the committed reference vectors, and the
    16x112x112 output shape with the final block's missing ReLU (negative values survive)."""
    import os
    from dualvar_b200 import backbones as PB
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backbones_next.npz"))
    ref, prod = _pair("r2d3d18")
    x = torch.randn(2, 3, 4, 64, 64, generator=torch.Generator().manual_seed(5)).to(dev)
    with torch.no_grad():
        yt = prod(x)                                  # train mode: batch statistics + running-stat update
        assert _rel(yt, torch.from_numpy(g["r2d3d18_out"]).to(dev)) < 6e-2     # 17 bf16 conv layers, 2x2 final map
        prod.eval()
        ye = prod(x)
        assert _rel(ye, torch.from_numpy(g["r2d3d18_out_eval"]).to(dev)) < 6e-2
        net, param = PB.select_backbone("r2d3d18")
        y = net.to(dev).train()(torch.randn(2, 3, 16, 112, 112, device=dev))
    assert param["feature_size"] == 256 and y.shape == (2, 256, 16, 4, 4) and torch.isfinite(y).all()
    assert bool((y < 0).any())


def test_s3dg_full_size_shape():
    from dualvar_b200 import backbones as PB
    net, param = PB.select_backbone("s3dg")
    net = net.to(dev).train()
    with torch.no_grad():
        y = net(torch.randn(2, 3, 32, 128, 128, device=dev))
    assert param["feature_size"] == 1024 and y.shape == (2, 1024, 4, 4, 4) and torch.isfinite(y).all()
