"""ORACLE (test infrastructure): restatement of the reference's contrastive objectives.

Written from the semantics in SURVEY.md Appendix C with explicit index arithmetic instead of the
reference's boolean-mask gathers; numerically the same matmul / cross-entropy calls in fp32.
Reference: model/simclr.py:183-337, model/moco.py:404-480, utils/utils.py:75-92,321-338,
classifier.py:963-983 (paths under /root/reference).
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F


# ----------------------------------------------------------------------------- gather
class GatherLayer(torch.autograd.Function):
    """all_gather whose backward keeps only this rank's slice, no reduce (utils/utils.py:321-338)."""

    @staticmethod
    def forward(ctx, x):
        out = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
        dist.all_gather(out, x.contiguous())
        return tuple(out)

    @staticmethod
    def backward(ctx, *grads):
        return grads[dist.get_rank()].clone()


def gather_cat(x, distributed):
    if not distributed:
        return x, 0, 1
    return torch.cat(GatherLayer.apply(x), dim=0), dist.get_rank(), dist.get_world_size()


# ----------------------------------------------------------------------------- NT-Xent / tc
def _positive_first_columns(row_ids, n_cols, n_half):
    """For each global row id r (in a view-major 2N ordering) build the reference column order:
    [positive, every other column except r itself in ascending order] (model/simclr.py:205-216)."""
    rows = row_ids.numel()
    pos = (row_ids + n_half) % n_cols
    cols = torch.arange(n_cols, device=row_ids.device).unsqueeze(0).expand(rows, n_cols)
    keep = (cols != row_ids.unsqueeze(1)) & (cols != pos.unsqueeze(1))
    neg = cols[keep].view(rows, n_cols - 2)
    return torch.cat([pos.unsqueeze(1), neg], dim=1)


def nt_xent(features, temperature, distributed=False):
    """Clip-level NT-Xent over the (gathered) global batch (model/simclr.py:183-229).
    features: (B, 2, d) unit-norm. Returns (logits (2N, 2N-1), labels, loss)."""
    B, V, d = features.shape
    assert V == 2
    feats, _, _ = gather_cat(features, distributed)
    N = feats.shape[0]
    f = feats.permute(1, 0, 2).reshape(2 * N, d)          # view-major rows r = v*N + i
    sim = f @ f.t()
    rows = torch.arange(2 * N, device=f.device)
    logits = sim.gather(1, _positive_first_columns(rows, 2 * N, N)) / temperature
    labels = torch.zeros(2 * N, dtype=torch.long, device=f.device)
    return logits, labels, F.cross_entropy(logits, labels)


def tc_loss(series, temperature, distributed=False):
    """Temporal-coherent inter-variant loss (model/simclr.py:280-337). series: (B, 2, s, e) with each
    e-vector unit-norm. Rows = this rank's 2B clips, columns = all 2N clips; similarity = mean over
    the s x s segment pairs. Returns (logits (2B, 2N-1), labels, loss)."""
    B, V, s, e = series.shape
    feats, rank, world = gather_cat(series, distributed)
    N = feats.shape[0]
    n = N // world
    col = feats.permute(1, 0, 2, 3).reshape(2 * N, s, e)
    row = feats[rank * n:(rank + 1) * n].permute(1, 0, 2, 3).reshape(2 * n, s, e)
    sim = torch.einsum("rse,cte->rcst", row, col).mean(dim=(2, 3))
    local = torch.arange(n, device=series.device) + rank * n
    row_ids = torch.cat([local, local + N])               # global view-major ids of the local rows
    logits = sim.gather(1, _positive_first_columns(row_ids, 2 * N, N)) / temperature
    labels = torch.zeros(2 * n, dtype=torch.long, device=series.device)
    return logits, labels, F.cross_entropy(logits, labels)


# ----------------------------------------------------------------------------- shuffle-rank
def rank_loss(pairs, theta, weight, clip_max=5.0):
    """Shuffle-rank intra-variant loss. pairs: (B, s, 2, e) (segment, view). For row (v,k) of the
    per-sample 2s x 2s Gram matrix the "highest" entry is the same segment in the other view and the
    "second" entries are all columns that are neither self nor highest, ascending.
    SimCLR: softplus(min(diff/theta, 5)) (model/simclr.py:231-278); MoCo: clip_max=None, theta=0.05
    (model/moco.py:440-480). Returns (margin_logits (B*2s, 2s-1), labels, loss)."""
    B, s, V, e = pairs.shape
    assert V == 2
    x = pairs.permute(0, 2, 1, 3).reshape(B, 2 * s, e)    # rows [view0 seg0..s-1, view1 seg0..s-1]
    gram = torch.bmm(x, x.transpose(1, 2))
    r = torch.arange(2 * s, device=pairs.device)
    order = _positive_first_columns(r, 2 * s, s)          # same structure: partner = (r + s) mod 2s
    picked = gram.gather(2, order.unsqueeze(0).expand(B, -1, -1))
    highest, second = picked[:, :, :1], picked[:, :, 1:]
    z = (second - highest) / theta
    if clip_max is not None:
        z = z.clip(max=clip_max)
    loss = weight * torch.log(1 + torch.exp(z)).mean()
    logits = picked.reshape(-1, 2 * s - 1)
    labels = torch.zeros(logits.shape[0], dtype=torch.long, device=pairs.device)
    return logits, labels, loss


# ----------------------------------------------------------------------------- segment shuffle
def draw_segment_perms(batch, n_series):
    """One np.random.permutation(n_series) per sample, in batch order, from the global NumPy stream
    (model/simclr.py:379-381, model/moco.py:544-546)."""
    return np.array([np.random.permutation(n_series) for _ in range(batch)], dtype=np.int64)


def shuffle_segments(clips, perms):
    """clips (B,C,T,H,W); shuffled segment j = original segment perms[b, j] (model/simclr.py:378-383)."""
    B, C, T, H, W = clips.shape
    s = perms.shape[1]
    seg = clips.reshape(B, C, s, T // s, H, W)
    idx = torch.as_tensor(perms, device=clips.device).view(B, 1, s, 1, 1, 1).expand_as(seg)
    return torch.gather(seg, 2, idx).reshape(B, C, T, H, W)


def calibrate_segments(series, perms):
    """series (B, s, e) of a shuffled clip: chunk j goes back to slot perms[b, j] (model/simclr.py:389-392)."""
    idx = torch.as_tensor(perms, device=series.device).unsqueeze(-1).expand_as(series)
    return torch.scatter(series, 1, idx, series)


# ----------------------------------------------------------------------------- MoCo
def moco_infonce(q, k, queue, temperature):
    """logits = [q.k, q.queue] / T, CE vs label 0 (model/moco.py:426-438). queue: (d, K)."""
    pos = (q * k).sum(dim=1, keepdim=True)
    neg = q @ queue.clone().detach()
    logits = torch.cat([pos, neg], dim=1) / temperature
    labels = torch.zeros(q.shape[0], dtype=torch.long, device=q.device)
    return logits, labels, F.cross_entropy(logits, labels)


def moco_tc(q, k, queue, temperature):
    """Series version: similarity = mean over s x s segment pairs (model/moco.py:404-424).
    q, k: (B, s, e); queue: (s*e, K)."""
    B, s, e = q.shape
    K = queue.shape[1]
    neg_feats = queue.clone().detach().t().reshape(K, s, e)
    pos = torch.einsum("bse,bte->bst", q, k).mean(dim=(1, 2)).unsqueeze(1)
    neg = torch.einsum("bse,kte->bkst", q, neg_feats).mean(dim=(2, 3))
    logits = torch.cat([pos, neg], dim=1) / temperature
    labels = torch.zeros(B, dtype=torch.long, device=q.device)
    return logits, labels, F.cross_entropy(logits, labels)


@torch.no_grad()
def momentum_update(params_q, params_k, m):
    """theta_k = m*theta_k + (1-m)*theta_q (model/moco.py:328-334)."""
    for pq, pk in zip(params_q, params_k):
        pk.data = pk.data * m + pq.data * (1.0 - m)


@torch.no_grad()
def enqueue(queue, ptr, keys):
    """queue[:, ptr:ptr+B] = keys.T ; returns the advanced pointer (model/moco.py:343-355)."""
    K = queue.shape[1]
    bs = keys.shape[0]
    assert K % bs == 0
    queue[:, ptr:ptr + bs] = keys.t()
    return (ptr + bs) % K


# ----------------------------------------------------------------------------- metrics / retrieval
def topk_accuracy(output, target, topk=(1,)):
    """Fraction of rows whose target is within the k best logits (utils/utils.py:75-92)."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    hit = pred.t().eq(target.view(1, -1))
    return [hit[:k].reshape(-1).float().sum(0) * (1.0 / target.size(0)) for k in topk]


def retrieval_topk(test_feat, train_feat, ks=(1, 5, 10, 20, 50)):
    """Centre each set by its own mean, L2-normalise, sim = test @ train.T, top-k indices
    (classifier.py:963-980). Returns (sim, {k: indices (n_test, k)})."""
    te = test_feat - test_feat.mean(dim=0, keepdim=True)
    tr = train_feat - train_feat.mean(dim=0, keepdim=True)
    te = F.normalize(te, p=2, dim=1)
    tr = F.normalize(tr, p=2, dim=1)
    sim = te @ tr.t()
    return sim, {k: torch.topk(sim, k, dim=1)[1] for k in ks}


def retrieval_accuracy(topk_idx, train_label, test_label):
    """hit if any of the k retrieved train labels equals the test label (classifier.py:981-983)."""
    return {k: (train_label[idx] == test_label.unsqueeze(1)).any(dim=1).float().mean().item()
            for k, idx in topk_idx.items()}
