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
