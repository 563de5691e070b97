"""Video retrieval on pooled backbone features (classifier.py:963-983): centre, L2-normalise,
sim = test @ train^T, top-k for k in (1, 5, 10, 20, 50), hit if any retrieved train label matches.

The chain is evaluated in fp64 from the fp32 features (csrc/retrieval.cu) so that the top-k indices are
reproducible bit for bit; see DESIGN.md "Retrieval" for how this relates to the reference's fp32 matmul.
"""
import torch

from . import _lib
from ._lib import ptr, stream_ptr

KS = (1, 5, 10, 20, 50)


def retrieval_topk(test_feature, train_feature, ks=KS, return_sim=True):
    """Returns (sim fp32 (n_test, n_train) or None, {k: int64 indices (n_test, k)})."""
    for t in (test_feature, train_feature):
        if not t.is_cuda:
            raise _lib.DualVarNativeError("retrieval runs on a B200 only (no CPU fallback)")
    te, tr = test_feature.contiguous().float(), train_feature.contiguous().float()
    nt, d = te.shape
    ntr = tr.shape[0]
    dev = te.device
    kmax = min(max(ks), ntr)
    te64 = torch.empty((nt, d), dtype=torch.float64, device=dev)
    tr64 = torch.empty((ntr, d), dtype=torch.float64, device=dev)
    mean = torch.empty(d, dtype=torch.float64, device=dev)
    _lib.call("dv_retrieval_prepare", ptr(te), ptr(mean), ptr(te64), nt, d, stream_ptr())
    _lib.call("dv_retrieval_prepare", ptr(tr), ptr(mean), ptr(tr64), ntr, d, stream_ptr())
    sim64 = torch.empty((nt, ntr), dtype=torch.float64, device=dev)
    sim32 = torch.empty((nt, ntr), dtype=torch.float32, device=dev) if return_sim else None
    idx = torch.empty((nt, kmax), dtype=torch.int64, device=dev)
    _lib.call("dv_retrieval_sim_topk", ptr(te64), ptr(tr64), ptr(sim64), ptr(sim32), ptr(idx), nt, ntr, d, kmax,
              stream_ptr())
    return sim32, {k: idx[:, :min(k, kmax)] for k in ks}


def retrieval_accuracy(topk_idx, train_label, test_label):
    """classifier.py:981-983."""
    return {k: (train_label[idx] == test_label.unsqueeze(1)).any(dim=1).float().mean().item()
            for k, idx in topk_idx.items()}
