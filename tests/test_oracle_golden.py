"""CPU: pin oracle/ (the torch restatement) to the fixtures generated from the real reference by
tests/golden/make_golden.py. Tolerances are fp32 round-off only: same ATen kernels, different
indexing path."""
import os
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import models as OM
from oracle import objectives as O
from oracle.backbones import select_backbone


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _seed(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_nt_xent_matches_reference(golden_dir):
    g = _load(golden_dir, "objectives.npz")
    f = _t(g["ntx_in"]).requires_grad_(True)
    logits, labels, loss = O.nt_xent(f, 0.07)
    loss.backward()
    assert logits.shape == (12, 11) and int(labels.sum()) == 0
    np.testing.assert_allclose(logits.detach().numpy(), g["ntx_logits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(loss.item(), g["ntx_loss"], rtol=1e-6)
    np.testing.assert_allclose(f.grad.numpy(), g["ntx_grad"], rtol=1e-4, atol=1e-6)


def test_tc_matches_reference(golden_dir):
    g = _load(golden_dir, "objectives.npz")
    f = _t(g["tc_in"]).requires_grad_(True)
    logits, _, loss = O.tc_loss(f, 0.07)
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), g["tc_logits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(loss.item(), g["tc_loss"], rtol=1e-6)
    np.testing.assert_allclose(f.grad.numpy(), g["tc_grad"], rtol=1e-4, atol=1e-6)


def test_rank_loss_matches_reference(golden_dir):
    g = _load(golden_dir, "objectives.npz")
    p = _t(g["rank_in"]).requires_grad_(True)
    logits, _, loss = O.rank_loss(p, 0.05, 0.5)
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), g["rank_logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(loss.item(), g["rank_loss"], rtol=1e-6)
    np.testing.assert_allclose(p.grad.numpy(), g["rank_grad"], rtol=1e-4, atol=1e-6)
    logits3, _, loss3 = O.rank_loss(_t(g["rank3_in"]), 0.05, 0.5)
    np.testing.assert_allclose(logits3.numpy(), g["rank3_logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(loss3.item(), g["rank3_loss"], rtol=1e-6)
    mlogits, _, mloss = O.rank_loss(_t(g["rank_in"]), 0.05, 0.5, clip_max=None)
    np.testing.assert_allclose(mlogits.numpy(), g["mrank_logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mloss.item(), g["mrank_loss"], rtol=1e-6)


def test_moco_losses_match_reference(golden_dir):
    g = _load(golden_dir, "objectives.npz")
    q = _t(g["moco_q"]).requires_grad_(True)
    logits, _, loss = O.moco_infonce(q, _t(g["moco_k"]), _t(g["moco_queue"]), 0.07)
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), g["moco_logits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(loss.item(), g["moco_loss"], rtol=1e-6)
    np.testing.assert_allclose(q.grad.numpy(), g["moco_grad"], rtol=1e-4, atol=1e-6)
    sq = _t(g["mtc_q"]).requires_grad_(True)
    logits, _, loss = O.moco_tc(sq, _t(g["mtc_k"]), _t(g["mtc_queue"]), 0.07)
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), g["mtc_logits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(loss.item(), g["mtc_loss"], rtol=1e-6)
    np.testing.assert_allclose(sq.grad.numpy(), g["mtc_grad"], rtol=1e-4, atol=1e-6)


def test_topk_and_retrieval_match_reference(golden_dir):
    g = _load(golden_dir, "objectives.npz")
    acc = O.topk_accuracy(_t(g["topk_in"]), torch.zeros(9, dtype=torch.long), (1, 5))
    np.testing.assert_allclose([a.item() for a in acc], g["topk_acc"], rtol=1e-6)
    sim, idx = O.retrieval_topk(_t(g["ret_test"]), _t(g["ret_train"]))
    np.testing.assert_allclose(sim.numpy(), g["ret_sim"], rtol=1e-5, atol=1e-6)
    for k in (1, 5, 10, 20, 50):
        assert np.array_equal(idx[k].numpy(), g[f"ret_top{k}"])


@pytest.mark.parametrize("name", ["r21d", "r3d", "c3d", "s3d", "s3dg"])
def test_backbone_matches_reference(golden_dir, name):
    g = _load(golden_dir, "backbones.npz")
    _seed(0)
    net, param = select_backbone(name)
    assert param["feature_size"] == int(g[f"{name}_feature_size"])
    assert sum(p.numel() for p in net.parameters()) == int(g[f"{name}_nparams"])
    assert sorted(net.state_dict().keys()) == list(g[f"{name}_keys"])
    checksum = float(sum(p.detach().double().abs().sum() for p in net.parameters()))
    np.testing.assert_allclose(checksum, float(g[f"{name}_checksum"]), rtol=1e-12)
    x = torch.randn(2, 3, 8, 32, 32, generator=torch.Generator().manual_seed(5))
    net.train()
    y = net(x)
    np.testing.assert_allclose(y.detach().numpy(), g[f"{name}_out"], rtol=1e-4, atol=1e-5)
    net.eval()
    with torch.no_grad():
        np.testing.assert_allclose(net(x).numpy(), g[f"{name}_out_eval"], rtol=1e-4, atol=1e-5)


def test_r2d3d18_matches_reference(golden_dir):
    """the oracle's r2d3d18 against vectors from the real backbone/resnet_2d3d.py."""
    g = _load(golden_dir, "backbones_next.npz")
    _seed(0)
    net, param = select_backbone("r2d3d18")
    assert param["feature_size"] == int(g["r2d3d18_feature_size"])
    assert sum(p.numel() for p in net.parameters()) == int(g["r2d3d18_nparams"])
    assert sorted(net.state_dict().keys()) == list(g["r2d3d18_keys"])
    checksum = float(sum(p.detach().double().abs().sum() for p in net.parameters()))
    np.testing.assert_allclose(checksum, float(g["r2d3d18_checksum"]), rtol=1e-12)
    x = torch.randn(2, 3, 4, 64, 64, generator=torch.Generator().manual_seed(5))
    net.train()
    np.testing.assert_allclose(net(x).detach().numpy(), g["r2d3d18_out"], rtol=1e-4, atol=1e-5)
    net.eval()
    with torch.no_grad():
        np.testing.assert_allclose(net(x).numpy(), g["r2d3d18_out_eval"], rtol=1e-4, atol=1e-5)
        assert tuple(net(torch.zeros(1, 3, 16, 112, 112)).shape) == tuple(g["r2d3d18_shape112"]) == (1, 256, 16, 4, 4)


def test_select_backbone_unknown_raises():
    with pytest.raises(NotImplementedError):
        select_backbone("resnet18")


@pytest.mark.parametrize("net", ["r3d", "r21d"])
def test_simclr_dualvar_step_matches_reference(golden_dir, net):
    g = _load(golden_dir, "steps.npz")
    _seed(0)
    m = OM.SimCLR_TimeSeriesV4(net, dim=128, T=0.07, distributed=False, n_series=2, series_dim=64,
                               aligned_T=0.07, mode="clip-sr-tc", args=SimpleNamespace(shufflerank_theta=0.05))
    m.train()
    x = torch.randn(2, 3, 3, 8, 32, 32, generator=torch.Generator().manual_seed(3))
    np.random.seed(7)
    ret = m(x)
    loss = sum(v for k, v in ret.items() if "loss" in k)
    loss.backward()
    for k, v in ret.items():
        np.testing.assert_allclose(v.detach().numpy(), g[f"simclr_{net}_{k}"], rtol=2e-4, atol=2e-5, err_msg=k)
    np.testing.assert_allclose(loss.item(), g[f"simclr_{net}_total"], rtol=1e-5)
    norms = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
    for n, ref in zip(g[f"simclr_{net}_gradnames"], g[f"simclr_{net}_gradnorms"]):
        np.testing.assert_allclose(norms[str(n)], ref, rtol=5e-3, atol=1e-7, err_msg=str(n))


def test_moco_dualvar_step_matches_reference(golden_dir):
    g = _load(golden_dir, "steps.npz")
    _seed(0)
    m = OM.MoCo_TimeSeriesV4("r21d", dim=128, K=16, m=0.9, T=0.07, distributed=False, n_series=2,
                             series_dim=64, aligned_T=0.07, mode="clip-sr-tc",
                             args=SimpleNamespace(shufflerank_theta=0.05))
    m.train()
    x = torch.randn(4, 3, 3, 8, 32, 32, generator=torch.Generator().manual_seed(4))
    np.random.seed(9)
    ret = m(x)
    loss = sum(v for k, v in ret.items() if "loss" in k)
    loss.backward()
    for k, v in ret.items():
        np.testing.assert_allclose(v.detach().numpy(), g[f"moco_{k}"], rtol=2e-4, atol=2e-5, err_msg=k)
    np.testing.assert_allclose(m.queue.numpy(), g["moco_queue_after"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(m.series_queue.numpy(), g["moco_series_queue_after"], rtol=1e-4, atol=1e-6)
    assert int(m.queue_ptr) == int(g["moco_ptr_after"][0])
    ck = float(sum(p.detach().double().abs().sum() for p in m.encoder_k.parameters()))
    np.testing.assert_allclose(ck, float(g["moco_kparam_checksum"]), rtol=1e-9)
    norms = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
    for n, ref in zip(g["moco_gradnames"], g["moco_gradnorms"]):
        np.testing.assert_allclose(norms[str(n)], ref, rtol=5e-3, atol=1e-7, err_msg=str(n))
