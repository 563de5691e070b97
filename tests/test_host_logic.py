"""CPU: host-side mirror of the reference interface — constructor signatures, state_dict keys and
initial weights of the drop-in modules equal the reference's (via the golden fixtures)."""
import inspect
import os
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from dualvar_b200 import backbones as PB
from dualvar_b200 import models as PM
from dualvar_b200.engine import RawClips
from oracle import models as OM


def _seed(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


@pytest.mark.parametrize("name", ["r21d", "r3d", "c3d", "s3d", "s3dg"])
def test_backbone_state_dict_and_init_match_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "backbones.npz"))
    _seed(0)
    net, param = PB.select_backbone(name)
    assert param["feature_size"] == int(g[f"{name}_feature_size"])
    assert sorted(net.state_dict().keys()) == list(g[f"{name}_keys"])
    assert sum(p.numel() for p in net.parameters()) == int(g[f"{name}_nparams"])
    checksum = float(sum(p.detach().double().abs().sum() for p in net.parameters()))
    np.testing.assert_allclose(checksum, float(g[f"{name}_checksum"]), rtol=1e-12)


def test_r2d3d18_state_dict_and_init_match_reference(golden_dir):
    """select_backbone('r2d3d18') — keys, parameter count, kaiming-normal(fan_out) init checksum."""
    g = np.load(os.path.join(golden_dir, "backbones_next.npz"))
    _seed(0)
    net, param = PB.select_backbone("r2d3d18")
    assert param["feature_size"] == int(g["r2d3d18_feature_size"]) == 256
    assert sorted(net.state_dict().keys()) == list(g["r2d3d18_keys"])
    assert sum(p.numel() for p in net.parameters()) == int(g["r2d3d18_nparams"]) == 5210176
    checksum = float(sum(p.detach().double().abs().sum() for p in net.parameters()))
    np.testing.assert_allclose(checksum, float(g["r2d3d18_checksum"]), rtol=1e-12)


def test_select_backbone_contract():
    with pytest.raises(NotImplementedError):
        PB.select_backbone("nope")
    sig = inspect.signature(PB.select_backbone)
    assert list(sig.parameters) == ["network", "first_channel"]


def test_true_18_layer_variant_is_available():
    net = PB.R2Plus1DNet((2, 2, 2, 2))
    assert sum(p.numel() for p in net.parameters()) == 33178423


@pytest.mark.parametrize("cls", ["SimCLR_TimeSeriesV4", "SimCLR_Naked", "LinearClassifier", "MoCo_Naked",
                                 "MoCo_TimeSeriesV4"])
def test_model_signatures_match_oracle(cls):
    a = inspect.signature(getattr(PM, cls).__init__)
    b = inspect.signature(getattr(OM, cls).__init__)
    assert list(a.parameters) == list(b.parameters)
    for k in a.parameters:
        assert a.parameters[k].default == b.parameters[k].default, k


def test_simclr_state_dict_interchanges_with_oracle():
    args = SimpleNamespace(shufflerank_theta=0.05)
    _seed(0)
    ref = OM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    _seed(0)
    prod = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    sr, sp = ref.state_dict(), prod.state_dict()
    assert list(sr.keys()) == list(sp.keys()) and len(sr) == 152
    for k in sr:
        assert sr[k].shape == sp[k].shape and torch.equal(sr[k], sp[k]), k
    prod.load_state_dict(sr)
    ref.load_state_dict(sp)
    assert sum(p.numel() for p in prod.parameters()) == 15021943


def test_sync_batchnorm_conversion_keeps_structure():
    prod = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                                  SimpleNamespace(shufflerank_theta=0.05))
    keys = list(prod.state_dict().keys())
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(prod)      # pretrain.py:244
    assert list(conv.state_dict().keys()) == keys
    assert sum(isinstance(m, torch.nn.SyncBatchNorm) for m in conv.modules()) == 24


def test_rawclips_layout():
    rc = RawClips(torch.zeros(2, 3, 48, 8, 8), 3)
    assert rc.block_shape == (2, 3, 3, 16, 8, 8)
    with pytest.raises(AssertionError):
        RawClips(torch.zeros(2, 3, 47, 8, 8), 3)


def test_moco_state_dict_interchanges_with_oracle():
    args = SimpleNamespace(shufflerank_theta=0.05)
    _seed(0)
    ref = OM.MoCo_TimeSeriesV4("r21d", 128, 64, 0.999, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    _seed(0)
    prod = PM.MoCo_TimeSeriesV4("r21d", 128, 64, 0.999, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    sr, sp = ref.state_dict(), prod.state_dict()
    assert list(sr.keys()) == list(sp.keys()) and len(sr) == 307
    assert all(torch.equal(sr[k], sp[k]) for k in sr)
    assert not any(p.requires_grad for p in prod.encoder_k.parameters())
    assert sum(p.numel() for p in prod.parameters() if p.requires_grad) == 15021943


def test_precision_switch_and_plane_products():
    """fp32 mode host logic: the switch validates its arguments, and a convolution over K split planes is the sum of
    the products (i, j) with i + j < K - every term above the 2^-8K truncation, smallest contributions first."""
    import dualvar_b200
    from dualvar_b200 import engine as E
    assert E.PRECISION == "bf16" and not E.fp32_mode()
    try:
        dualvar_b200.set_precision("fp32")
        assert E.fp32_mode() and E.F32_PLANES == 3
        t3 = E._terms()
        assert sorted(t3) == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)] and t3[-1] == (0, 0)
        assert all(a[0] + a[1] >= b[0] + b[1] for a, b in zip(t3, t3[1:]))
        dualvar_b200.set_precision("fp32", planes=2)
        assert sorted(E._terms()) == [(0, 0), (0, 1), (1, 0)]
        with pytest.raises(ValueError):
            dualvar_b200.set_precision("fp16")
        with pytest.raises(ValueError):
            dualvar_b200.set_precision("fp32", planes=4)
    finally:
        dualvar_b200.set_precision("bf16", planes=3)
    assert not E.fp32_mode()


def test_split_planes_identity_in_numpy():
    """The split the fp32 mode rests on, restated in numpy: p_k = bf16(x - p_0 - .. - p_(k-1)) has exact residuals,
    three planes reproduce every normal fp32 value bit for bit, two planes leave at most 2^-16 relative."""
    import numpy as np

    def bf16(a):
        u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000          # round to nearest even on the top 16 bits
        return u.astype(np.uint32).view(np.float32)

    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1 << 16) * np.logspace(-3, 3, 1 << 16)).astype(np.float32)
    p0 = bf16(x); r1 = x - p0
    p1 = bf16(r1); r2 = r1 - p1
    p2 = bf16(r2)
    assert np.array_equal((p0.astype(np.float64) + p1 + p2).astype(np.float32), x)
    assert np.array_equal(p0 + p1 + p2, x)
    assert np.all(np.abs((p0.astype(np.float64) + p1) - x) <= np.abs(x) * 2.0 ** -16)
