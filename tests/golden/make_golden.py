"""Generate the golden fixtures that pin ``oracle/`` to the real reference.

Runs ONLY in the build container (needs /root/reference). It imports the UNMODIFIED reference
modules with the shim of SURVEY.md Appendix B (two stub modules, the ``calc_contrast_loss`` alias the
reference forgot, and ``Tensor.cuda`` -> identity on this CPU-only box), feeds them seeded inputs and
stores the outputs as small .npz files next to this script. ``tests/test_oracle_golden.py`` then
checks the oracle against these files without needing the reference.

    python tests/golden/make_golden.py
"""
import os
import random
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = os.environ.get("DUALVAR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name, attr in (("IPython", "embed"), ("dataloader", "KVReader")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            setattr(m, attr, None)
            sys.modules[name] = m
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import model as ref_model  # noqa
    from backbone.select_backbone import select_backbone as ref_select  # noqa
    import utils.utils as ref_utils  # noqa
    ref_model.SimCLR_TimeSeriesV4.calc_contrast_loss = ref_model.SimCLR_TimeSeriesV4.calc_clip_contrast_loss
    ref_model.MoCo_TimeSeriesV4.calc_contrast_loss = ref_model.MoCo_TimeSeriesV4.calc_clip_contrast_loss
    return ref_model, ref_select, ref_utils


def seed_all(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def npy(t):
    return t.detach().cpu().numpy()


def unit(t, dim=-1):
    return torch.nn.functional.normalize(t, dim=dim)


def golden_objectives(ref_model, ref_utils):
    out = {}
    args = SimpleNamespace(shufflerank_theta=0.05)
    seed_all(0)
    sim = ref_model.SimCLR_TimeSeriesV4("r3d", dim=16, T=0.07, distributed=False, n_series=2, series_dim=8,
                                        aligned_T=0.07, args=args)
    g = torch.Generator().manual_seed(11)
    # clip NT-Xent
    f = unit(torch.randn(6, 2, 16, generator=g)).requires_grad_(True)
    r = sim.calc_clip_contrast_loss(f, 2)
    r["clip_contrast_loss"].backward()
    out.update(ntx_in=npy(f), ntx_logits=npy(r["clip_logits"]), ntx_loss=npy(r["clip_contrast_loss"]),
               ntx_grad=npy(f.grad))
    # tc
    sfeat = unit(torch.randn(5, 2, 2, 8, generator=g)).requires_grad_(True)
    r = sim.calc_tc_contrast_loss(sfeat)
    r["tc_contrast_loss"].backward()
    out.update(tc_in=npy(sfeat), tc_logits=npy(r["tc_logits"]), tc_loss=npy(r["tc_contrast_loss"]),
               tc_grad=npy(sfeat.grad))
    # shuffle-rank (SimCLR: clip at 5, theta from args)
    pairs = unit(torch.randn(4, 2, 2, 8, generator=g)).requires_grad_(True)
    r = sim.calc_ranking_loss(pairs, 2, "x_", weight=0.5)
    r["x_margin_contrast_loss"].backward()
    out.update(rank_in=npy(pairs), rank_logits=npy(r["x_margin_logits"]), rank_loss=npy(r["x_margin_contrast_loss"]),
               rank_grad=npy(pairs.grad))
    # three segments
    sim3 = ref_model.SimCLR_TimeSeriesV4("r3d", dim=16, T=0.07, distributed=False, n_series=3, series_dim=8, args=args)
    pairs3 = unit(torch.randn(3, 3, 2, 8, generator=g))
    r = sim3.calc_ranking_loss(pairs3, 2, "x_", weight=0.5)
    out.update(rank3_in=npy(pairs3), rank3_logits=npy(r["x_margin_logits"]), rank3_loss=npy(r["x_margin_contrast_loss"]))
    # MoCo variants
    moco = ref_model.MoCo_TimeSeriesV4("r3d", dim=16, K=32, m=0.99, T=0.07, distributed=False, n_series=2,
                                       series_dim=8, aligned_T=0.07, args=args)
    q = unit(torch.randn(4, 16, generator=g)).requires_grad_(True)
    k = unit(torch.randn(4, 16, generator=g))
    queue = unit(torch.randn(16, 32, generator=g), dim=0)
    r = moco.calc_clip_contrast_loss(q, k, queue)
    r["clip_contrast_loss"].backward()
    out.update(moco_q=npy(q), moco_k=npy(k), moco_queue=npy(queue), moco_logits=npy(r["clip_logits"]),
               moco_loss=npy(r["clip_contrast_loss"]), moco_grad=npy(q.grad))
    sq = unit(torch.randn(4, 2, 8, generator=g)).requires_grad_(True)
    sk = unit(torch.randn(4, 2, 8, generator=g))
    squeue = unit(torch.randn(2, 8, 32, generator=g), dim=1).reshape(16, 32)
    r = moco.calc_tc_contrast_loss(sq, sk, squeue)
    r["tc_contrast_loss"].backward()
    out.update(mtc_q=npy(sq), mtc_k=npy(sk), mtc_queue=npy(squeue), mtc_logits=npy(r["tc_logits"]),
               mtc_loss=npy(r["tc_contrast_loss"]), mtc_grad=npy(sq.grad))
    r = moco.calc_ranking_loss(pairs.detach(), 2, "x_", weight=0.5)
    out.update(mrank_logits=npy(r["x_margin_logits"]), mrank_loss=npy(r["x_margin_contrast_loss"]))
    # top-k accuracy
    logits = torch.randn(9, 13, generator=g)
    target = torch.zeros(9, dtype=torch.long)
    acc = ref_utils.calc_topk_accuracy(logits, target, (1, 5))
    out.update(topk_in=npy(logits), topk_acc=np.array([a.item() for a in acc]))
    # retrieval maths, restated from classifier.py:963-983 (the driver script itself cannot be imported)
    te = torch.randn(37, 32, generator=g)
    tr = torch.randn(53, 32, generator=g)
    te_n = torch.nn.functional.normalize(te - te.mean(dim=0, keepdim=True), p=2, dim=1)
    tr_n = torch.nn.functional.normalize(tr - tr.mean(dim=0, keepdim=True), p=2, dim=1)
    simm = torch.matmul(te_n, tr_n.t())
    out.update(ret_test=npy(te), ret_train=npy(tr), ret_sim=npy(simm))
    for kk in (1, 5, 10, 20, 50):
        out[f"ret_top{kk}"] = npy(torch.topk(simm, kk, dim=1)[1])
    np.savez_compressed(os.path.join(HERE, "objectives.npz"), **out)
    print("objectives.npz", len(out), "arrays")


def golden_backbones(ref_select):
    out = {}
    for name in ("r21d", "r3d", "c3d", "s3d", "s3dg"):
        seed_all(0)
        net, param = ref_select(name)
        n_params = sum(p.numel() for p in net.parameters())
        checksum = float(sum(p.detach().double().abs().sum() for p in net.parameters()))
        keys = sorted(net.state_dict().keys())
        x = torch.randn(2, 3, 8, 32, 32, generator=torch.Generator().manual_seed(5))
        net.train()
        y = net(x)
        out[f"{name}_nparams"] = np.array(n_params)
        out[f"{name}_checksum"] = np.array(checksum)
        out[f"{name}_feature_size"] = np.array(param["feature_size"])
        out[f"{name}_out"] = npy(y)
        out[f"{name}_keys"] = np.array(keys)
        # a second forward in eval mode exercises the running statistics updated by the first
        net.eval()
        with torch.no_grad():
            out[f"{name}_out_eval"] = npy(net(x))
        with torch.no_grad():
            z = net(torch.zeros(1, 3, 16, 112, 112))
        out[f"{name}_shape112"] = np.array(z.shape)
        print(name, n_params, tuple(y.shape), tuple(z.shape))
    np.savez_compressed(os.path.join(HERE, "backbones.npz"), **out)


def golden_backbones_next(ref_select):
    """Backbones outside the north-star list, added as SURVEY §8(f) rows: written to their own fixture file."""
    out = {}
    for name in ("r2d3d18",):
        seed_all(0)
        net, param = ref_select(name)
        out[f"{name}_nparams"] = np.array(sum(p.numel() for p in net.parameters()))
        out[f"{name}_checksum"] = np.array(float(sum(p.detach().double().abs().sum() for p in net.parameters())))
        out[f"{name}_feature_size"] = np.array(param["feature_size"])
        out[f"{name}_keys"] = np.array(sorted(net.state_dict().keys()))
        x = torch.randn(2, 3, 4, 64, 64, generator=torch.Generator().manual_seed(5))
        net.train()
        y = net(x)
        out[f"{name}_out"] = npy(y)
        net.eval()
        with torch.no_grad():
            out[f"{name}_out_eval"] = npy(net(x))
            out[f"{name}_shape112"] = np.array(net(torch.zeros(1, 3, 16, 112, 112)).shape)
        print(name, int(out[f"{name}_nparams"]), tuple(y.shape))
    np.savez_compressed(os.path.join(HERE, "backbones_next.npz"), **out)


def grads_digest(model):
    d = {}
    for n, p in model.named_parameters():
        if p.grad is not None:
            d[n] = float(p.grad.double().norm())
    return d


def golden_steps(ref_model):
    out = {}
    args = SimpleNamespace(shufflerank_theta=0.05)
    for net in ("r3d", "r21d"):
        seed_all(0)
        m = ref_model.SimCLR_TimeSeriesV4(net, dim=128, T=0.07, distributed=False, n_series=2, series_dim=64,
                                          aligned_T=0.07, mode="clip-sr-tc", args=args)
        m.train()
        x = torch.randn(2, 3, 3, 8, 32, 32, generator=torch.Generator().manual_seed(3))
        np.random.seed(7)
        perms = np.array([np.random.permutation(2) for _ in range(2)])
        np.random.seed(7)
        ret = m(x)
        loss = sum(v for k, v in ret.items() if "loss" in k)
        loss.backward()
        out[f"simclr_{net}_perms"] = perms
        for k, v in ret.items():
            out[f"simclr_{net}_{k}"] = npy(v)
        out[f"simclr_{net}_total"] = npy(loss)
        gd = grads_digest(m)
        out[f"simclr_{net}_gradnames"] = np.array(list(gd.keys()))
        out[f"simclr_{net}_gradnorms"] = np.array(list(gd.values()))
        print("simclr", net, float(loss))
    seed_all(0)
    m = ref_model.MoCo_TimeSeriesV4("r21d", dim=128, K=16, m=0.9, T=0.07, distributed=False, n_series=2,
                                    series_dim=64, aligned_T=0.07, mode="clip-sr-tc", args=args)
    m.train()
    x = torch.randn(4, 3, 3, 8, 32, 32, generator=torch.Generator().manual_seed(4))
    np.random.seed(9)
    perms = np.array([np.random.permutation(2) for _ in range(4)])
    np.random.seed(9)
    ret = m(x)
    loss = sum(v for k, v in ret.items() if "loss" in k)
    loss.backward()
    out["moco_perms"] = perms
    for k, v in ret.items():
        out[f"moco_{k}"] = npy(v)
    out["moco_total"] = npy(loss)
    out["moco_queue_after"] = npy(m.queue)
    out["moco_series_queue_after"] = npy(m.series_queue)
    out["moco_ptr_after"] = npy(m.queue_ptr)
    out["moco_kparam_checksum"] = np.array(float(sum(p.detach().double().abs().sum() for p in m.encoder_k.parameters())))
    gd = grads_digest(m)
    out["moco_gradnames"] = np.array(list(gd.keys()))
    out["moco_gradnorms"] = np.array(list(gd.values()))
    print("moco", float(loss))
    np.savez_compressed(os.path.join(HERE, "steps.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref_model, ref_select, ref_utils = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "next":      # only the SURVEY §8(f) additions (backbones_next.npz)
        golden_backbones_next(ref_select)
        sys.exit(0)
    golden_objectives(ref_model, ref_utils)
    golden_backbones(ref_select)
    golden_backbones_next(ref_select)
    golden_steps(ref_model)
