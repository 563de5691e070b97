"""Evaluation drivers of classifier.py restated on the drop-in modules.

* ``temporal_10clip_eval``  - classifier.py:657-738: every video contributes 10 temporally uniform clips; the clips'
  softmax probabilities are averaged per video, top-1/top-5 per video (``summarize_probability``, classifier.py:762-784).
* ``extract_video_features`` - classifier.py:873-903: pooled backbone features of the 10 clips, averaged per video.
* ``retrieval_eval``         - classifier.py:963-983 on such features (centre, normalise, similarity, k-NN hit rates)
  through dualvar_b200.retrieval on a GPU.

A loader batch is ``{'seq': (B, 3, 10*seq_len, H, W) frames in [0,1] (or uint8), 'vid': labels[, 'vpath'/'vname']}`` as
the reference's 10-clip datasets yield it; ``Normalize`` + the view/permute of ``tr()`` (classifier.py:673-680) are the
ingest kernel's job when the model is a dualvar_b200 module (engine.RawClips with 10 "views" of one clip each).
"""
import torch
import torch.nn.functional as F

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def clips_of(frames, num_seq=10, native=True, mean=MEAN, std=STD):
    """Loader frames (B, 3, num_seq*seq_len, H, W) -> what the model takes: (B*num_seq, 3, seq_len, H, W) normalised
    (classifier.py:673-680). For the dualvar_b200 modules on CUDA the whole transform is folded into the ingest."""
    B, C, L, H, W = frames.shape
    assert L % num_seq == 0
    if native and frames.is_cuda:
        from .engine import RawClips
        return RawClips(frames, num_seq, mean=mean, std=std)
    x = frames.float() / 255.0 if frames.dtype == torch.uint8 else frames.float()
    m = torch.tensor(mean, device=x.device).view(1, C, 1, 1, 1)
    s = torch.tensor(std, device=x.device).view(1, C, 1, 1, 1)
    x = (x - m) / s
    return x.view(B, C, num_seq, L // num_seq, H, W).permute(0, 2, 1, 3, 4, 5).contiguous().view(B * num_seq, C, L // num_seq, H, W)


def _forward(model, frames, num_seq, native):
    x = clips_of(frames, num_seq, native)
    if native and frames.is_cuda:
        # LinearClassifier.forward takes the 5-D clip batch or RawClips (clips = views, in (b, view) order)
        return model(x)
    return model(x)


@torch.no_grad()
def temporal_10clip_eval(model, loader, num_seq=10, native=True, device=None):
    """classifier.py:657-738. Returns {'top1', 'top5', 'n_videos', 'mean_prob' (n_videos, num_class), 'label'}."""
    model.eval()
    probs, labels = [], []
    for batch in loader:
        frames = batch["seq"] if device is None else batch["seq"].to(device, non_blocking=True)
        B = frames.shape[0]
        logit, _ = _forward(model, frames, num_seq, native)
        prob = F.softmax(logit.float(), dim=-1).view(B, num_seq, -1).mean(1)          # average over the temporal window
        probs.append(prob)
        labels.append(torch.as_tensor(batch["vid"]).view(-1).to(prob.device))
    prob = torch.cat(probs)
    label = torch.cat(labels).long()
    k5 = min(5, prob.shape[1])
    top = prob.topk(k5, dim=1)[1]
    hit = top == label.unsqueeze(1)
    return {"top1": hit[:, :1].any(1).float().mean().item() * 100.0, "top5": hit.any(1).float().mean().item() * 100.0,
            "n_videos": int(prob.shape[0]), "mean_prob": prob, "label": label}


@torch.no_grad()
def extract_video_features(model, loader, num_seq=10, native=True, device=None):
    """classifier.py:873-903: (video feature = mean over the clips' pooled features, per-clip features, labels)."""
    model.eval()
    feats, per, labels = [], [], []
    for batch in loader:
        frames = batch["seq"] if device is None else batch["seq"].to(device, non_blocking=True)
        B = frames.shape[0]
        _, feature = _forward(model, frames, num_seq, native)
        pf = feature.float().view(B, num_seq, -1)
        per.append(pf)
        feats.append(pf.mean(dim=1))
        labels.append(torch.as_tensor(batch["vid"]).view(-1).to(pf.device))
    return torch.cat(feats), torch.cat(per), torch.cat(labels).long()


def retrieval_eval(test_feature, test_label, train_feature, train_label):
    """classifier.py:963-983 on the GPU kernels: {k: k-NN hit rate} for k in (1, 5, 10, 20, 50)."""
    from . import retrieval as R
    _, idx = R.retrieval_topk(test_feature, train_feature, return_sim=False)
    return R.retrieval_accuracy(idx, train_label, test_label)
