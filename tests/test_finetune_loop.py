"""Finetune / linear-probe driver (classifier.py:390-654): host logic on CPU with the oracle classifier
against a literal restatement of the reference loop; on the GPU the product classifier, the cross-entropy kernel and the
ten-crop test transform against torch / Pillow."""
import copy
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dualvar_b200 import finetune_loop as FL


def _seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


def _loader(n_batches, B, T, H, n_class=7, device="cpu", seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"seq": torch.rand(B, 3, T, H, H, generator=g).to(device),
             "vid": torch.randint(0, n_class, (B,), generator=g).to(device)} for _ in range(n_batches)]


def _reference_epoch(model, loader, optimizer, train_what):
    """classifier.py:435-476 literally (CPU): mode, tr(), CE, zero_grad / backward / step."""
    model.eval() if train_what == "last" else model.train()
    mean = torch.tensor(FL.MEAN).view(1, 3, 1, 1, 1); std = torch.tensor(FL.STD).view(1, 3, 1, 1, 1)
    ce = torch.nn.CrossEntropyLoss()
    losses = []
    for batch in loader:
        x = batch["seq"]
        B, _, L, H, W = x.shape
        x = ((x - mean) / std).view(B, 3, 1, L, H, W).transpose(1, 2).contiguous().squeeze(1)
        logit, _ = model(x)
        loss = ce(logit, batch["vid"])
        losses.append((loss.item(), B))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
    return sum(l * b for l, b in losses) / sum(b for _, b in losses)


@pytest.mark.parametrize("train_what", ["last", "all"])
def test_train_epoch_matches_reference_loop_cpu(train_what):
    from oracle import models as OM
    _seed(0)
    a = OM.LinearClassifier(num_class=7, network="r3d", use_dropout=False)
    b = copy.deepcopy(a)
    loader = _loader(2, 2, 4, 32)
    opt_a = FL.build_optimizer(a, train_what, "sgd", lr=0.1, wd=1e-3)
    # the reference's parameter selection (classifier.py:233-247)
    params = []
    for name, p in b.named_parameters():
        if train_what == "last" and "backbone" in name:
            p.requires_grad = False
        else:
            params.append({"params": p})
    opt_b = torch.optim.SGD(params, lr=0.1, weight_decay=1e-3, momentum=0.9)
    assert len(opt_a.param_groups) == len(opt_b.param_groups) == (2 if train_what == "last" else len(list(a.parameters())))
    m = FL.train_one_epoch(loader, a, opt_a, train_what, native=False)
    ref_loss = _reference_epoch(b, loader, opt_b, train_what)
    assert abs(m["loss"] - ref_loss) < 1e-6 and m["n"] == 4
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(pa, pb), n
    if train_what == "last":       # frozen BatchNorm: eval mode leaves the running statistics untouched
        assert int(a.backbone.bn1.num_batches_tracked) == 0
    v = FL.validate(loader, a, native=False)
    assert 0.0 <= v["top1"] <= v["top5"] <= 100.0 and not a.training


def test_lr_schedule_and_fit_checkpoints_cpu(tmp_path):
    from oracle import models as OM
    _seed(1)
    model = OM.LinearClassifier(num_class=5, network="r3d", use_dropout=True, use_final_bn=True)
    opt = FL.build_optimizer(model, "last", "sgd", lr=1.0, wd=0.0)
    FL.adjust_learning_rate(opt, 3, (3, 5))
    assert all(abs(g["lr"] - 0.1) < 1e-12 for g in opt.param_groups)
    FL.adjust_learning_rate(opt, 4, (3, 5))
    assert all(abs(g["lr"] - 0.1) < 1e-12 for g in opt.param_groups)
    loader = _loader(1, 2, 4, 32, n_class=5)
    hist, best = FL.fit(model, loader, loader, opt, epochs=2, schedule=(1,), train_what="last", use_bn=True, eval_freq=1,
                        model_path=str(tmp_path), native=False, log=lambda *_: None)
    assert len(hist) == 2 and "val" in hist[1] and (tmp_path / "latest.pth.tar").exists() and (tmp_path / "epoch1.pth.tar").exists()
    ck = torch.load(tmp_path / "latest.pth.tar", weights_only=False)
    assert set(ck) == {"epoch", "state_dict", "best_acc", "optimizer", "iteration"} and ck["epoch"] == 1
    assert int(model.final_bn.num_batches_tracked) == 2       # final_bn trains while the backbone stays frozen


def test_five_crop_offsets_follow_reference_including_quirk():
    # A.FiveCrop on a 171 x 128 (w x h) frame with size 112 (utils/augmentation.py:203-220)
    w, h, t = 171, 128, 112
    assert FL.five_crop_offsets(w, h, t, t, 1) == (0, 0)
    assert FL.five_crop_offsets(w, h, t, t, 2) == (w - t, 0)
    assert FL.five_crop_offsets(w, h, t, t, 3) == (0, h - t)
    assert FL.five_crop_offsets(w, h, t, t, 4) == (w - t, h - t)          # tw == th here; the quirk shows when they differ
    assert FL.five_crop_offsets(200, 100, 60, 40, 4) == (140, 40)         # top edge = h - tw (as written in the reference)
    assert FL.five_crop_offsets(w, h, t, t, 5) == (int(round((w - t) / 2.)), int(round((h - t) / 2.)))
    with pytest.raises(ValueError):
        FL.five_crop_offsets(100, 100, 112, 112, 5)


def test_ten_crop_protocol_summary_cpu():
    """test_10crop's bookkeeping with a stub model: 10 augmentations per video, every one scored separately."""
    class Stub(torch.nn.Module):
        def forward(self, x):
            # class = rounded mean brightness of the clip -> depends on the loader's augmentation
            m = x.mean(dim=(1, 2, 3, 4))
            logit = torch.stack([-(m - c) ** 2 * 50 for c in range(3)], 1)
            return logit, m
    calls = []

    def make_loader(aug, flip):
        calls.append((aug, flip))
        base = torch.zeros(2, 3, 8, 4, 4)
        # video "a" (label 1) is right for every augmentation; video "b" (label 2) only when flipped
        val = {"a": 1.0, "b": 2.0 if flip else 0.0}
        mean = torch.tensor(FL.MEAN).view(1, 3, 1, 1, 1); std = torch.tensor(FL.STD).view(1, 3, 1, 1, 1)
        seq = torch.stack([base[0] + val["a"], base[1] + val["b"]]) * std + mean
        return [{"seq": seq, "vid": torch.tensor([1, 2]), "vpath": ["a", "b"]}]

    out = FL.test_10crop(make_loader, Stub(), mode="ten", seq_len=4, native=False)
    assert calls == [(a, f) for f in (0, 1) for a in (5, 1, 2, 3, 4)]
    assert out["center"] == (50.0, 100.0) and out["five"] == (50.0, 100.0)
    assert out["ten"] == (75.0, 100.0)        # video b: 5 of its 10 rows are right


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cross_entropy_kernel_matches_torch():
    from dualvar_b200 import objectives as O
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(3)
    for B, C in ((6, 101), (33, 51), (1, 7)):
        logit = (torch.randn(B, C, device=dev, generator=g) * 3).requires_grad_(True)
        target = torch.randint(0, C, (B,), device=dev, generator=g)
        loss, hits = O.cross_entropy(logit, target)
        (2.5 * loss).backward()
        ref = logit.detach().clone().requires_grad_(True)
        lr = F.cross_entropy(ref, target)
        (2.5 * lr).backward()
        assert abs(loss.item() - lr.item()) <= 1e-5 * max(1.0, abs(lr.item()))
        assert torch.allclose(logit.grad, ref.grad, rtol=1e-4, atol=1e-6)
        top = ref.topk(min(5, C), dim=1)[1]
        assert int(hits[0]) == int((top[:, 0] == target).sum()) and int(hits[1]) == int((top == target[:, None]).any(1).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("train_what", ["last", "all"])
def test_finetune_step_gradients_match_oracle(train_what, measured):
    """One finetune iteration of the product classifier against the oracle from identical weights: loss, every trained
    parameter's gradient tensor by tensor against its own bf16 noise floor (fp32 oracle vs the oracle with the product's
    rounding points, tests/bf16_emulation.py), and the frozen / trained split of train_what."""
    from bf16_emulation import compare_to_noise_floor, emulate_bf16, round_conv_weights
    from dualvar_b200 import models as PM
    from oracle import models as OM
    dev = "cuda:0"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _seed(0)
    kw = dict(num_class=101, network="r21d", use_dropout=False, use_final_bn=True)
    ref = round_conv_weights(OM.LinearClassifier(**kw).to(dev))
    with torch.no_grad():           # give the frozen BatchNorm layers non-trivial running statistics
        ref.train()
        m = torch.tensor(FL.MEAN, device=dev).view(1, 3, 1, 1, 1); s = torch.tensor(FL.STD, device=dev).view(1, 3, 1, 1, 1)
        ref((torch.rand(16, 3, 8, 64, 64, device=dev) - m) / s)
    emu = emulate_bf16(ref)
    prod = PM.LinearClassifier(**kw)
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev)
    loader = _loader(1, 32, 8, 64, n_class=101, device=dev, seed=4)
    loader[0]["seq"] = loader[0]["seq"].bfloat16().float()
    res = {}
    for name, model in (("ref", ref), ("emu", emu), ("prod", prod)):
        opt = FL.build_optimizer(model, train_what, "sgd", lr=0.0, wd=0.0)      # lr 0: gradients only, weights stay equal
        if name == "prod":
            assert type(opt).__module__.startswith("dualvar_b200")
        res[name] = FL.train_one_epoch(loader, model, opt, train_what, use_bn=True, native=(name == "prod"))
    measured(f"finetune_{train_what}.loss_rel", abs(res["prod"]["loss"] - res["ref"]["loss"]) / res["ref"]["loss"])
    assert abs(res["prod"]["loss"] - res["ref"]["loss"]) <= 1e-2 * res["ref"]["loss"]
    rows = compare_to_noise_floor(ref, emu, prod)
    worst = max(rows, key=lambda t: t[1] / (t[2] + 0.03))
    measured(f"finetune_{train_what}.grad_worst_tensor", {"name": worst[0], "err_product": worst[1], "err_rounding_oracle": worst[2]})
    assert len(rows) == (4 if train_what == "last" else len(list(ref.parameters())))
    bad = [(n, round(p, 4), round(e, 4)) for n, p, e in rows if p > 1.6 * e + 0.03]
    assert not bad, bad
    if train_what == "last":
        assert all(p.grad is None for n, p in prod.named_parameters() if "backbone" in n)
        assert int(prod.backbone.bn1.num_batches_tracked) == int(ref.backbone.bn1.num_batches_tracked)   # frozen BN


@pytest.mark.gpu
def test_ten_crop_clips_bit_exact_with_pillow():
    """flip + Scale((128,171)) bicubic + FiveCrop + ToTensor on the GPU == the reference's PIL chain
    (classifier.py:593-603, utils/augmentation.py:125-220,314-331)."""
    from PIL import Image
    dev = "cuda:0"
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (2, 4, 60, 80, 3), dtype=np.uint8)
    for where in (1, 2, 3, 4, 5):
        for flip in (0, 1):
            got = FL.ten_crop_clips(torch.from_numpy(frames).to(dev), where, flip, n_views=1).cpu().numpy()   # (B, 3, F, 112, 112)
            for b in range(2):
                for f in range(4):
                    im = Image.fromarray(frames[b, f])
                    if flip:
                        im = im.transpose(Image.FLIP_LEFT_RIGHT)
                    im = im.resize((128, 171), Image.BICUBIC)
                    left, top = FL.five_crop_offsets(128, 171, 112, 112, where)
                    want = np.asarray(im.crop((left, top, left + 112, top + 112))).transpose(2, 0, 1)
                    assert np.array_equal(got[b, :, f], want), (where, flip, b, f)
