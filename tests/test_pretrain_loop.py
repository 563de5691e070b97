"""Host logic of the pretraining driver on CPU with the oracle model; GPU: two epochs of the product
model, checkpoint, resume, and the resumed run continues bit for bit."""
import os
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from dualvar_b200 import pretrain_loop as L


def _seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


def test_total_loss_follows_reference_order():
    ret = {"clip_logits": torch.zeros(2, 3), "clip_labels": torch.zeros(2, dtype=torch.long),
           "clip_contrast_loss": torch.tensor(1.0), "tc_contrast_loss": torch.tensor(2.0),
           "aug_ranking_margin_contrast_loss": torch.tensor(4.0), "misc_loss": torch.tensor(8.0)}
    assert float(L.total_loss(ret)) == 15.0
    assert float(L.total_loss({"tc_contrast_loss": torch.tensor(2.0)})) == 2.0


def test_checkpoint_files_and_resume_cpu(tmp_path):
    from oracle import models as OM
    _seed(0)
    args = SimpleNamespace(shufflerank_theta=0.05)
    model = OM.SimCLR_TimeSeriesV4("r3d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    opt = L.build_optimizer(model, lr=0.01)
    assert isinstance(opt, torch.optim.SGD) and len(opt.param_groups) == len(list(model.parameters()))
    data = [{"seq": torch.randn(2, 3, 3, 4, 32, 32)} for _ in range(2)]
    hist, best, it = L.fit(model, data, opt, epochs=2, schedule=[1], model_path=str(tmp_path), log=lambda *_: None)
    assert len(hist) == 2 and it == 5 and {"loss", "clip_contrast_loss", "clip_acc", "tc_acc"} <= set(hist[0])
    assert abs(opt.param_groups[0]["lr"] - 0.001) < 1e-12            # MultiStepLR gamma 0.1 after epoch 1
    files = sorted(os.listdir(tmp_path))
    # pretrain.py:354 calls save_checkpoint(gap=0): the "previous epoch" it prunes is the file it is about to write, so
    # every epoch's file stays (utils/utils.py:19-24)
    assert {"latest.pth.tar", "epoch0.pth.tar", "epoch1.pth.tar"} <= set(files)
    ckpt = torch.load(tmp_path / "latest.pth.tar", weights_only=False)
    assert set(ckpt) == {"epoch", "state_dict", "best_acc", "optimizer", "iteration"} and ckpt["epoch"] == 1
    model2 = OM.SimCLR_TimeSeriesV4("r3d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    opt2 = L.build_optimizer(model2, lr=0.01)
    start, best2, it2 = L.load_checkpoint(str(tmp_path / "latest.pth.tar"), model2, opt2)
    assert start == 2 and it2 == 5
    for (n, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), n


@pytest.mark.gpu
def test_product_model_trains_checkpoints_and_resumes(tmp_path):
    from dualvar_b200 import models as PM
    from dualvar_b200.engine import RawClips
    dev = "cuda:0"
    args = SimpleNamespace(shufflerank_theta=0.05)

    def make():
        _seed(0)
        m = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args).to(dev)
        return m, L.build_optimizer(m, lr=0.01)

    gen = torch.Generator().manual_seed(3)
    data = [torch.rand(4, 3, 24, 64, 64, generator=gen).to(dev) for _ in range(3)]   # loader frames (B,3,3*T,H,W)
    to_input = lambda x: RawClips(x, 3)                                                # noqa: E731
    model, opt = make()
    np.random.seed(11)
    L.fit(model, data, opt, epochs=1, model_path=str(tmp_path), to_input=to_input, log=lambda *_: None)
    np.random.seed(12)
    h_a, _, _ = L.fit(model, data, opt, epochs=2, start_epoch=1, to_input=to_input, log=lambda *_: None)
    model2, opt2 = make()
    start, _, it = L.load_checkpoint(str(tmp_path / "latest.pth.tar"), model2, opt2, map_location=dev)
    assert start == 1 and it == 4
    # the restored state is the saved one bit for bit: weights, BN buffers and the fused optimizer's momentum buffers
    ckpt = torch.load(tmp_path / "latest.pth.tar", map_location=dev, weights_only=False)
    for k, v in model2.state_dict().items():
        assert torch.equal(v, ckpt["state_dict"][k]), k
    saved_bufs = [st["momentum_buffer"] for st in ckpt["optimizer"]["state"].values()]
    new_bufs = [opt2.state[p]["momentum_buffer"] for g in opt2.param_groups for p in g["params"]]
    assert len(saved_bufs) == len(new_bufs) and all(torch.equal(a, b) for a, b in zip(saved_bufs, new_bufs))
    np.random.seed(12)
    h_b, _, _ = L.fit(model2, data, opt2, epochs=2, start_epoch=1, iteration=it, to_input=to_input, log=lambda *_: None)
    # the continued and the resumed epoch see the same data and permutations; bf16 kernels with atomic reductions are not
    # run-to-run deterministic, so the averaged losses agree to a few parts in a thousand, not bit for bit
    assert np.isfinite(h_a[0]["loss"]) and abs(h_a[0]["loss"] - h_b[0]["loss"]) <= 3e-2 * abs(h_a[0]["loss"])


@pytest.mark.gpu
def test_fit_with_graphed_step_and_lr_schedule():
    """fit(..., graph_step=GraphedTrainStep): epochs of CUDA-graph replays, MultiStepLR milestones re-capture the graph,
    meters and checkpoint state behave like the eager loop."""
    from dualvar_b200 import models as PM
    from dualvar_b200.graph_step import GraphedTrainStep
    dev = "cuda:0"
    _seed(0)
    m = PM.SimCLR_TimeSeriesV4("r3d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                               SimpleNamespace(shufflerank_theta=0.05)).to(dev)
    opt = L.build_optimizer(m, lr=1e-3)
    gen = torch.Generator().manual_seed(3)
    data = [torch.rand(4, 3, 24, 32, 32, generator=gen).pin_memory() for _ in range(4)]
    step = GraphedTrainStep(m, opt, n_views=3, warmup=1)
    hist, _, it = L.fit(m, data, opt, epochs=3, schedule=(1,), graph_step=step, log=lambda *_: None)
    assert it == 13 and len(hist) == 3 and all(np.isfinite(h["loss"]) for h in hist)
    assert step.eager_steps == 1 and step.captures == 2 and step.replays == 11        # lr changed once after epoch 0
    assert abs(opt.param_groups[0]["lr"] - 1e-4) < 1e-12
    assert {"loss", "clip_contrast_loss", "clip_acc", "tc_contrast_loss"} <= set(hist[0])
    assert int(m.encoder_q[0].bn1.num_batches_tracked) == 24
