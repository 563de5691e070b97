"""fp32 heads and contrastive objectives as small autograd Functions over the C-ABI kernels.

These operate on tiny tensors (<= 2N x 2N similarities, N = global batch), so torch autograd is
used to chain them; each Function's arithmetic is kernels from csrc/heads.cu and csrc/losses.cu.
Gradients of the losses are produced in the same pass as the loss (the row kernel overwrites the
similarity matrix with dLoss/dS) and only scaled by the incoming gradient in backward.

Reference semantics: model/simclr.py:183-337, model/moco.py:404-480,
utils/utils.py:321-338 (GatherLayer: all_gather forward, local slice of the gradient backward).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr, stream_ptr

call = _lib.call
_f = ctypes.c_float


def _dist_on(distributed):
    return bool(distributed) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _sgemm(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias=None, relu=0):
    call("dv_sgemm", ta, tb, M, N, K, _f(alpha), ptr(A), lda, ptr(B), ldb, _f(beta), ptr(C), ldc, ptr(bias),
         relu, stream_ptr())


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (+ReLU); W is an nn.Conv3d 1x1x1 weight (N, K, 1, 1, 1) or a Linear weight (N, K).
    Reference: nn.Conv3d(.., kernel_size=1) heads, model/simclr.py:168-180."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x = x.contiguous()
        M, K = x.shape
        N = weight.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        _sgemm(0, 1, M, N, K, 1.0, x, K, weight, K, 0.0, y, N, bias, 1 if relu else 0)
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.relu, ctx.has_bias = relu, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        M, K = x.shape
        N = weight.shape[0]
        dy = dy.contiguous()
        if ctx.relu:
            d = torch.empty_like(dy)
            call("dv_relu_bwd", ptr(dy), ptr(y), ptr(d), dy.numel(), stream_ptr())
            dy = d
        dx = torch.empty_like(x)
        _sgemm(0, 0, M, K, N, 1.0, dy, N, weight, K, 0.0, dx, K)
        dw = torch.empty_like(weight)
        _sgemm(1, 0, N, K, M, 1.0, dy, N, x, K, 0.0, dw, K)
        db = None
        if ctx.has_bias:
            db = torch.empty(N, dtype=torch.float32, device=x.device)
            call("dv_colsum", ptr(dy), ptr(db), M, N, N, _f(0.0), stream_ptr())
        return dx, dw, db, None


def linear(x, mod, relu=False):
    return LinearFn.apply(x, mod.weight, mod.bias, relu)


class L2NormFn(torch.autograd.Function):
    """F.normalize over the last dimension (model/simclr.py:359,367,393)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        d = x.shape[-1]
        rows = x.numel() // d
        y = torch.empty_like(x)
        inv = torch.empty(rows, dtype=torch.float32, device=x.device)
        call("dv_l2norm_fwd", ptr(x), ptr(y), ptr(inv), rows, d, _f(1e-12), stream_ptr())
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = dy.contiguous()
        d = y.shape[-1]
        dx = torch.empty_like(y)
        call("dv_l2norm_bwd", ptr(dy), ptr(y), ptr(inv), ptr(dx), y.numel() // d, d, stream_ptr())
        return dx


def l2norm(x):
    return L2NormFn.apply(x)


class SegmentMeanFn(torch.autograd.Function):
    """(rows, s, e) -> (rows, e) mean over segments. The tc similarity, a mean over the s x s segment
    pairs of dot products (model/simclr.py:297,304), equals the dot product of these means."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        rows, s, e = x.shape
        out = torch.empty((rows, e), dtype=torch.float32, device=x.device)
        call("dv_segment_sum", ptr(x), ptr(out), rows, s, e, _f(1.0 / s), stream_ptr())
        ctx.s = s
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        rows, e = dout.shape
        dx = torch.empty((rows, ctx.s, e), dtype=torch.float32, device=dout.device)
        call("dv_segment_bcast", ptr(dout), ptr(dx), rows, ctx.s, e, _f(1.0 / ctx.s), _f(0.0), stream_ptr())
        return dx


class PermuteSegmentsFn(torch.autograd.Function):
    """out[b, perm[b, j]] = in[b, j]: re-align the series of a shuffled clip (model/simclr.py:389-392)."""

    @staticmethod
    def forward(ctx, x, perm):
        x = x.contiguous()
        B, s, e = x.shape
        out = torch.empty_like(x)
        call("dv_permute_segments", ptr(x), ptr(out), ptr(perm), B, s, e, 0, stream_ptr())
        ctx.perm = perm
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        B, s, e = dout.shape
        dx = torch.empty_like(dout)
        call("dv_permute_segments", ptr(dout), ptr(dx), ptr(ctx.perm), B, s, e, 1, stream_ptr())
        return dx, None


_index_cache = {}


def _sim_blocks(C):
    """Column blocks of the fused similarity / log-sum-exp launch (dv_sim_ce_blocks: 128 columns each)."""
    return (int(C) + 127) // 128


def _row_indices(n, N, rank, local_only, device):
    key = (n, N, rank, local_only, str(device))
    hit = _index_cache.get(key)
    if hit is None:
        if local_only:
            base = torch.arange(n, dtype=torch.int64) + rank * n
            rows = torch.cat([base, base + N])
        else:
            rows = torch.arange(2 * N, dtype=torch.int64)
        self_col = rows.to(torch.int32)
        pos_col = ((rows + N) % (2 * N)).to(torch.int32)
        hit = (rows.to(device), self_col.to(device), pos_col.to(device))
        _index_cache[key] = hit
    return hit


class ContrastFn(torch.autograd.Function):
    """NT-Xent over the (all-gathered) global batch with the reference's logits layout.

    feats: (n, 2, d) local unit vectors. local_rows=False: rows = all 2N clips (clip loss,
    model/simclr.py:183-229); local_rows=True: rows = this rank's 2n clips against all 2N columns
    (tc loss, model/simclr.py:280-337). Backward returns only the local slice of the gradient, as
    GatherLayer does (utils/utils.py:334-338). Returns (loss, logits, hits[top1, top5])."""

    @staticmethod
    def forward(ctx, feats, temperature, distributed, local_rows):
        feats = feats.contiguous()
        n, V, d = feats.shape
        dev = feats.device
        rank, world = 0, 1
        if _dist_on(distributed):
            rank, world = dist.get_rank(), dist.get_world_size()
            gathered = torch.empty((world * n, V, d), dtype=feats.dtype, device=dev)
            dist.all_gather_into_tensor(gathered, feats)
        else:
            gathered = feats
        N = gathered.shape[0]
        f_all = gathered.permute(1, 0, 2).contiguous().view(2 * N, d)      # view-major rows v*N + i
        rows, self_col, pos_col = _row_indices(n, N, rank, local_rows, dev)
        f_rows = f_all.index_select(0, rows) if local_rows else f_all
        R = f_rows.shape[0]
        # similarity GEMM fused with the row-wise log-sum-exp (csrc/sim_ce.cu): S, the logits in reference order and
        # the online-softmax partials in one launch; the finish launch turns S into dLoss/dS
        S = torch.empty((R, 2 * N), dtype=torch.float32, device=dev)
        logits = torch.empty((R, 2 * N - 1), dtype=torch.float32, device=dev)
        loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        hits = torch.zeros(2, dtype=torch.int32, device=dev)
        partials = torch.empty((R, _sim_blocks(2 * N), 2), dtype=torch.float32, device=dev)
        call("dv_sim_ce_fwd", ptr(f_rows), d, ptr(f_all), d, 0, R, 2 * N, d, ptr(S), 2 * N, 0, ptr(logits), 2 * N - 1,
             ptr(self_col), ptr(pos_col), _f(1.0 / temperature), ptr(partials), stream_ptr())
        call("dv_sim_ce_finish", ptr(S), 2 * N, R, 2 * N, 0, ptr(partials), ptr(logits), 2 * N - 1, ptr(self_col),
             ptr(pos_col), _f(1.0 / temperature), _f(1.0 / R), ptr(loss_sum), ptr(hits), stream_ptr())
        # S now holds dLoss/dS. Column-role gradient for every clip, row-role gradient for the rows.
        d_all = torch.empty((2 * N, d), dtype=torch.float32, device=dev)
        _sgemm(1, 0, 2 * N, d, R, 1.0, S, 2 * N, f_rows, d, 0.0, d_all, d)
        if local_rows:
            d_rows = torch.empty((R, d), dtype=torch.float32, device=dev)
            _sgemm(0, 0, R, d, 2 * N, 1.0, S, 2 * N, f_all, d, 0.0, d_rows, d)
            d_all.index_add_(0, rows, d_rows)
        else:
            _sgemm(0, 0, R, d, 2 * N, 1.0, S, 2 * N, f_all, d, 1.0, d_all, d)
        d_local = d_all.view(2, N, d)[:, rank * n:(rank + 1) * n].permute(1, 0, 2).contiguous()
        ctx.save_for_backward(d_local)
        ctx.mark_non_differentiable(logits, hits)
        return (loss_sum / R).squeeze(0), logits, hits

    @staticmethod
    def backward(ctx, dloss, _dlogits, _dhits):
        (d_local,) = ctx.saved_tensors
        return d_local * dloss, None, None, None


class RankLossFn(torch.autograd.Function):
    """Shuffle-rank loss on two aligned series a, b: (B, s, e) (model/simclr.py:231-278;
    clip_max=None -> MoCo variant model/moco.py:440-480). Returns (loss, margin_logits, hits)."""

    @staticmethod
    def forward(ctx, a, b, theta, weight, clip_max):
        a, b = a.contiguous(), b.contiguous()
        B, s, e = a.shape
        dev = a.device
        da, db = torch.empty_like(a), torch.empty_like(b)
        logits = torch.empty((B * 2 * s, 2 * s - 1), dtype=torch.float32, device=dev)
        loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        hits = torch.zeros(1, dtype=torch.int32, device=dev)
        call("dv_rank_loss", ptr(a), ptr(b), ptr(da), ptr(db), ptr(logits), ptr(loss_sum), ptr(hits), B, s, e,
             _f(theta), _f(clip_max if clip_max is not None else -1.0), _f(weight), stream_ptr())
        ctx.save_for_backward(da, db)
        ctx.mark_non_differentiable(logits, hits)
        return loss_sum.squeeze(0), logits, hits

    @staticmethod
    def backward(ctx, dloss, _dl, _dh):
        da, db = ctx.saved_tensors
        return da * dloss, db * dloss, None, None, None


class QueueContrastFn(torch.autograd.Function):
    """MoCo InfoNCE: logits = [q.k, q.queue] / T against label 0 (model/moco.py:426-438).
    queue: (d, K) buffer, keys k are constants. Returns (loss, logits, hits)."""

    @staticmethod
    def forward(ctx, q, k, queue, temperature):
        q, k = q.contiguous(), k.contiguous()
        B, d = q.shape
        K = queue.shape[1]
        dev = q.device
        # column 0 = q.k (the positive), columns 1..K = q.queue: the queue GEMM fused with the row-wise log-sum-exp
        # (csrc/sim_ce.cu) reads the (d, K) queue buffer as it is; the finish launch folds column 0 in
        S = torch.empty((B, K + 1), dtype=torch.float32, device=dev)
        call("dv_rowdot", ptr(q), ptr(k), ptr(S), B, d, K + 1, stream_ptr())
        logits = torch.empty((B, K + 1), dtype=torch.float32, device=dev)
        loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        hits = torch.zeros(2, dtype=torch.int32, device=dev)
        pos = torch.zeros(B, dtype=torch.int32, device=dev)
        partials = torch.empty((B, _sim_blocks(K), 2), dtype=torch.float32, device=dev)
        call("dv_sim_ce_fwd", ptr(q), d, ptr(queue), K, 1, B, K, d, ptr(S), K + 1, 1, ptr(logits), K + 1, None, ptr(pos),
             _f(1.0 / temperature), ptr(partials), stream_ptr())
        call("dv_sim_ce_finish", ptr(S), K + 1, B, K, 1, ptr(partials), ptr(logits), K + 1, None, ptr(pos),
             _f(1.0 / temperature), _f(1.0 / B), ptr(loss_sum), ptr(hits), stream_ptr())
        dq = torch.empty_like(q)
        _sgemm(0, 1, B, d, K, 1.0, S[:, 1:], K + 1, queue, K, 0.0, dq, d)
        call("dv_row_axpy", ptr(S), K + 1, ptr(k), ptr(dq), B, d, _f(1.0), stream_ptr())
        ctx.save_for_backward(dq)
        ctx.mark_non_differentiable(logits, hits)
        return (loss_sum / B).squeeze(0), logits, hits

    @staticmethod
    def backward(ctx, dloss, _dl, _dh):
        (dq,) = ctx.saved_tensors
        return dq * dloss, None, None, None


# ----------------------------------------------------------------------------- public helpers
def _zeros_labels(n, device):
    return torch.zeros(n, dtype=torch.long, device=device)


def nt_xent(features, temperature, distributed, prefix="clip_"):
    loss, logits, hits = ContrastFn.apply(features, temperature, distributed, False)
    return {f"{prefix}logits": logits, f"{prefix}labels": _zeros_labels(logits.shape[0], logits.device),
            f"{prefix}contrast_loss": loss}, hits


def tc_loss(series, temperature, distributed, prefix="tc_"):
    n, V, s, e = series.shape
    means = SegmentMeanFn.apply(series.reshape(n * V, s, e)).view(n, V, e)
    loss, logits, hits = ContrastFn.apply(means, temperature, distributed, True)
    return {f"{prefix}logits": logits, f"{prefix}labels": _zeros_labels(logits.shape[0], logits.device),
            f"{prefix}contrast_loss": loss}, hits


def rank_loss(a, b, theta, weight, clip_max, prefix):
    loss, logits, hits = RankLossFn.apply(a, b, theta, weight, clip_max)
    return {f"{prefix}margin_logits": logits, f"{prefix}margin_labels": _zeros_labels(logits.shape[0], logits.device),
            f"{prefix}margin_contrast_loss": loss}, hits


def queue_contrast(q, k, queue, temperature, prefix):
    loss, logits, hits = QueueContrastFn.apply(q, k, queue, temperature)
    return {f"{prefix}logits": logits, f"{prefix}labels": _zeros_labels(logits.shape[0], logits.device),
            f"{prefix}contrast_loss": loss}, hits


class CrossEntropyFn(torch.autograd.Function):
    """nn.CrossEntropyLoss (mean) + calc_topk_accuracy(logit, target, (1, 5)) of the finetune / linear-probe loop
    (classifier.py:465-467,525-527; utils/utils.py:75-92) in one launch of the contrastive row kernel: the "positive"
    column of row r is target[r], no column is dropped, T = 1; the kernel also leaves dLoss/dlogits behind.
    Returns (loss, hits[top1, top5])."""

    @staticmethod
    def forward(ctx, logits, target):
        B, C = logits.shape
        dev = logits.device
        S = logits.detach().float().contiguous().clone()
        scratch = torch.empty((B, C), dtype=torch.float32, device=dev)
        loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        hits = torch.zeros(2, dtype=torch.int32, device=dev)
        pos = target.to(torch.int32).contiguous()
        call("dv_contrast_rows", ptr(S), ptr(scratch), None, ptr(pos), B, C, C, C, _f(1.0), _f(1.0 / B), ptr(loss_sum),
             ptr(hits), stream_ptr())
        ctx.save_for_backward(S)
        ctx.mark_non_differentiable(hits)
        return (loss_sum / B).squeeze(0), hits

    @staticmethod
    def backward(ctx, dloss, _dh):
        (dS,) = ctx.saved_tensors
        return dS * dloss, None


def cross_entropy(logits, target):
    """(mean CE loss, int32 [top-1 hits, top-5 hits]) for (B, num_class) logits and int64 targets."""
    return CrossEntropyFn.apply(logits, target)
