"""Evaluation drivers: host logic on CPU with the oracle classifier against a literal restatement of the
reference loop; on the GPU the product classifier through the fused ingest against the oracle."""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dualvar_b200 import eval_loop as EL


def _seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


def _loader(n_batches, B, T, H, device="cpu", seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"seq": torch.rand(B, 3, 10 * T, H, H, generator=g).to(device),
             "vid": torch.randint(0, 7, (B,), generator=g)} for _ in range(n_batches)]


def _reference_10clip(model, loader):
    """classifier.py:673-738 literally: tr(), softmax, mean over the 10 clips, top-k per video."""
    model.eval()
    hits1 = hits5 = n = 0
    mean = torch.tensor(EL.MEAN).view(1, 3, 1, 1, 1); std = torch.tensor(EL.STD).view(1, 3, 1, 1, 1)
    with torch.no_grad():
        for batch in loader:
            x = batch["seq"]
            B, _, L, H, W = x.shape
            x = ((x - mean) / std).view(B, 3, 10, L // 10, H, W).permute(0, 2, 1, 3, 4, 5).contiguous().view(B * 10, 3, L // 10, H, W)
            logit, _ = model(x)
            prob = F.softmax(logit, dim=-1).view(B, 10, -1).mean(1)
            top5 = prob.topk(5, dim=1)[1]
            hits1 += (top5[:, 0] == batch["vid"]).sum().item()
            hits5 += (top5 == batch["vid"].unsqueeze(1)).any(1).sum().item()
            n += B
    return 100.0 * hits1 / n, 100.0 * hits5 / n


def test_10clip_eval_and_features_match_reference_loop_cpu():
    from oracle import models as OM
    _seed(0)
    model = OM.LinearClassifier(num_class=7, network="r3d", use_dropout=True)
    loader = _loader(2, 2, 4, 32)
    out = EL.temporal_10clip_eval(model, loader, native=False)
    t1, t5 = _reference_10clip(model, loader)
    assert out["n_videos"] == 4 and abs(out["top1"] - t1) < 1e-9 and abs(out["top5"] - t5) < 1e-9
    assert torch.allclose(out["mean_prob"].sum(1), torch.ones(4), atol=1e-5)
    feat, per, label = EL.extract_video_features(model, loader, native=False)
    assert feat.shape == (4, 512) and per.shape == (4, 10, 512) and label.shape == (4,)
    assert torch.allclose(feat, per.mean(1))


@pytest.mark.gpu
def test_10clip_eval_product_matches_oracle_and_retrieval_runs():
    from dualvar_b200 import models as PM
    from oracle import models as OM
    dev = "cuda:0"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _seed(0)
    ref = OM.LinearClassifier(num_class=7, network="r21d", use_dropout=True).to(dev)
    prod = PM.LinearClassifier(num_class=7, network="r21d", use_dropout=True)
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev)
    loader = _loader(2, 3, 8, 64, device=dev)
    a = EL.temporal_10clip_eval(prod, loader)                 # fused ingest (RawClips, 10 clips per video)
    b = EL.temporal_10clip_eval(ref, loader, native=False)
    assert a["n_videos"] == b["n_videos"] == 6
    assert (a["mean_prob"] - b["mean_prob"]).abs().max().item() < 2e-2      # probabilities, bf16 encoder
    # uint8 frames give the same result as their float image
    u8 = [{"seq": (x["seq"] * 255).round().to(torch.uint8), "vid": x["vid"]} for x in loader]
    f32 = [{"seq": x["seq"].float() / 255, "vid": x["vid"]} for x in u8]
    c, d = EL.temporal_10clip_eval(prod, u8), EL.temporal_10clip_eval(prod, f32)
    assert torch.equal(c["mean_prob"], d["mean_prob"])
    te, _, tl = EL.extract_video_features(prod, loader)
    tr_, _, trl = EL.extract_video_features(prod, _loader(3, 3, 8, 64, device=dev, seed=1))
    acc = EL.retrieval_eval(te, tl, tr_, trl)
    assert set(acc) == {1, 5, 10, 20, 50} and all(0.0 <= v <= 1.0 for v in acc.values())
    assert acc[50] >= acc[1]
