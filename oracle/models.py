"""ORACLE (test infrastructure): restatement of the reference's model wrappers — encoder + heads +
DualVar objectives — on top of oracle.backbones / oracle.objectives. Runs on CPU or GPU in fp32.

Reference: model/simclr.py:19-400, model/moco.py:28-573, model/classifier.py:9-84.
Differences from the reference as shipped are only the repairs SURVEY.md §0.3 lists as necessary to
run it at all: the forward calls the clip loss that exists (the reference calls a missing
``calc_contrast_loss`` alias) and label/index tensors are created on the input's device instead of
``.cuda()``.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import objectives as O
from .backbones import select_backbone


def _proj_head(cin, cout):
    return [nn.Conv3d(cin, cin, kernel_size=1, bias=True), nn.ReLU(), nn.Conv3d(cin, cout, kernel_size=1, bias=True)]


def _pack(prefix, logits, labels, loss, margin=False):
    mid = "margin_" if margin else ""
    return {f"{prefix}{mid}logits": logits, f"{prefix}{mid}labels": labels, f"{prefix}{mid}contrast_loss": loss}


class SimCLR_Naked(nn.Module):
    """Two-view SimCLR (model/simclr.py:19-121)."""

    def __init__(self, network="s3d", dim=128, T=0.07, distributed=True, nonlinear=True):
        super().__init__()
        self.dim, self.T, self.distributed, self.nonlinear = dim, T, distributed, nonlinear
        backbone, self.param = select_backbone(network)
        fs = self.param["feature_size"]
        self.encoder_q = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_q.extend(_proj_head(fs, dim))
        self.criterion = nn.CrossEntropyLoss()

    def forward(self, block):
        B, n_views = block.shape[:2]
        assert n_views == 2
        f = block.reshape(-1, *block.shape[2:])
        for mod in self.encoder_q:
            f = mod(f)
        f = F.normalize(f, dim=1).reshape(B, n_views, self.dim)
        return _pack("clip_", *O.nt_xent(f, self.T, self.distributed))


class SimCLR_TimeSeriesV4(nn.Module):
    """SimCLR + DualVar (model/simclr.py:130-400): clip NT-Xent on views 0,1; tc loss on the series
    head of views 0,1; two shuffle-rank losses against the segment-shuffled view 2."""

    def __init__(self, network="s3d", dim=128, T=0.07, distributed=True, nonlinear=True, n_series=2,
                 series_dim=64, series_T=0.07, aligned_T=0.07, mode="clip-sr-tc", args=None):
        super().__init__()
        self.args = args
        self.dim, self.T, self.distributed, self.nonlinear = dim, T, distributed, nonlinear
        self.n_series, self.series_dim = n_series, series_dim
        self.series_T, self.aligned_T, self.mode = series_T, aligned_T, mode
        self.with_clip, self.with_sr, self.with_tc = "clip" in mode, "sr" in mode, "tc" in mode
        backbone, self.param = select_backbone(network)
        fs = self.param["feature_size"]
        self.encoder_q = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear and self.with_clip:
            self.encoder_q.extend(_proj_head(fs, dim))
        self.criterion = nn.CrossEntropyLoss()
        self.series_proj_head = nn.Sequential(*_proj_head(fs, series_dim * n_series))

    def forward(self, block, perms=None):
        """block (B,3,C,T,H,W). ``perms`` (B, n_series) overrides the NumPy draw (tests only)."""
        block = block.contiguous()
        B, V, C, T, H, W = block.shape
        assert V == 3
        s, e = self.n_series, self.series_dim
        f = block.reshape(B * 3, C, T, H, W)
        pooled = None
        for i, mod in enumerate(self.encoder_q):
            f = mod(f)
            if i == 1:
                pooled = f
        f = F.normalize(f, dim=1).reshape(B, 3, self.dim)[:, :2].contiguous()
        ret = {}
        if self.with_clip:
            ret.update(_pack("clip_", *O.nt_xent(f, self.T, self.distributed)))
        series = F.normalize(self.series_proj_head(pooled).reshape(B, 3, s, e), dim=3)
        if self.with_tc:
            ret.update(_pack("tc_", *O.tc_loss(series[:, :2].contiguous(), self.aligned_T, self.distributed)))
        if self.with_sr:
            if perms is None:
                perms = O.draw_segment_perms(B, s)
            g = O.shuffle_segments(block[:, 2], perms)
            for mod in list(self.encoder_q)[:2]:
                g = mod(g)
            shuf = self.series_proj_head(g).reshape(B, s, e)
            shuf = F.normalize(O.calibrate_segments(shuf, perms), dim=2)
            theta = self.args.shufflerank_theta
            for view, prefix in ((0, "aug_ranking_"), (2, "unaug_ranking_")):
                pairs = torch.stack([series[:, view], shuf], dim=2).contiguous()   # (B, s, 2, e)
                ret.update(_pack(prefix, *O.rank_loss(pairs, theta, 0.5), margin=True))
        return ret


@torch.no_grad()
def _all_gather_cat(t):
    """model/moco.py:14-25."""
    out = [torch.ones_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t, async_op=False)
    return torch.cat(out, dim=0)


class _MoCoBase(nn.Module):
    @torch.no_grad()
    def _shuffle_ddp(self, x):
        """shuffle-BN: gather all key clips, rank 0's randperm is broadcast, each rank takes its
        slice (model/moco.py:357-383)."""
        n_this = x.shape[0]
        xg = _all_gather_cat(x)
        idx = torch.randperm(xg.shape[0]).to(x.device)
        dist.broadcast(idx, src=0)
        unshuf = torch.argsort(idx)
        mine = idx.view(xg.shape[0] // n_this, -1)[dist.get_rank()]
        return xg[mine], unshuf

    @torch.no_grad()
    def _unshuffle_ddp(self, x, unshuf):
        """model/moco.py:385-402."""
        n_this = x.shape[0]
        xg = _all_gather_cat(x)
        return xg[unshuf.view(xg.shape[0] // n_this, -1)[dist.get_rank()]]


class MoCo_Naked(_MoCoBase):
    """Two-view MoCo (model/moco.py:28-239)."""

    def __init__(self, network="s3d", dim=128, K=2048, m=0.999, T=0.07, distributed=True, nonlinear=True):
        super().__init__()
        self.dim, self.K, self.m, self.T = dim, K, m, T
        self.distributed, self.nonlinear = distributed, nonlinear
        backbone, self.param = select_backbone(network)
        fs = self.param["feature_size"]
        self.encoder_q = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_q.extend(_proj_head(fs, dim))
        backbone, _ = select_backbone(network)
        self.encoder_k = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_k.extend(_proj_head(fs, dim))
        for pq, pk in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            pk.data.copy_(pq.data)
            pk.requires_grad = False
        self.register_buffer("queue", torch.randn(dim, K))
        self.queue = F.normalize(self.queue, dim=0)
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
        self.criterion = nn.CrossEntropyLoss()

    def forward(self, block):
        B, N = block.shape[:2]
        assert N == 2
        x1, x2 = block[:, 0].contiguous(), block[:, 1].contiguous()
        q = x1
        for mod in self.encoder_q:
            q = mod(q)
        q = F.normalize(q, dim=1).reshape(B, self.dim)
        train = q.requires_grad
        with torch.no_grad():
            if train:
                O.momentum_update(self.encoder_q.parameters(), self.encoder_k.parameters(), self.m)
            if self.distributed:
                x2, unshuf = self._shuffle_ddp(x2)
            k = x2
            for mod in self.encoder_k:
                k = mod(k)
            k = F.normalize(k, dim=1)
            if self.distributed:
                k = self._unshuffle_ddp(k, unshuf)
        k = k.reshape(B, self.dim)
        ret = _pack("clip_", *O.moco_infonce(q, k, self.queue, self.T))
        if train:
            keys = _all_gather_cat(k) if self.distributed else k
            self.queue_ptr[0] = O.enqueue(self.queue, int(self.queue_ptr), keys)
        return ret


class MoCo_TimeSeriesV4(_MoCoBase):
    """MoCo + DualVar (model/moco.py:242-573)."""

    def __init__(self, network="s3d", dim=128, K=2048, m=0.999, T=0.07, distributed=True, nonlinear=True,
                 n_series=2, series_dim=64, series_T=0.07, aligned_T=0.07, mode="clip-sr-tc", args=None):
        super().__init__()
        self.dim, self.K, self.m, self.T = dim, K, m, T
        self.distributed, self.nonlinear = distributed, nonlinear
        self.n_series, self.series_dim, self.mode = n_series, series_dim, mode
        self.series_T, self.aligned_T = series_T, aligned_T
        self.with_clip, self.with_sr, self.with_tc = "clip" in mode, "sr" in mode, "tc" in mode
        backbone, self.param = select_backbone(network)
        fs = self.param["feature_size"]
        self.encoder_q = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_q.extend(_proj_head(fs, dim))
        self.series_proj_head_q = nn.Sequential(*_proj_head(fs, series_dim * n_series))
        backbone, _ = select_backbone(network)
        self.encoder_k = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_k.extend(_proj_head(fs, dim))
        self.series_proj_head_k = nn.Sequential(*_proj_head(fs, series_dim * n_series))
        for a, b in ((self.encoder_q, self.encoder_k), (self.series_proj_head_q, self.series_proj_head_k)):
            for pq, pk in zip(a.parameters(), b.parameters()):
                pk.data.copy_(pq.data)
                pk.requires_grad = False
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
        self.register_buffer("queue", torch.randn(dim, K))
        self.queue = F.normalize(self.queue, dim=0)
        self.register_buffer("series_queue", torch.randn(series_dim * n_series, K))
        self.series_queue = F.normalize(self.series_queue.view(n_series, series_dim, K), dim=1).view(
            n_series * series_dim, K)
        self.criterion = nn.CrossEntropyLoss()

    def _encode(self, encoder, head, x):
        pooled = None
        f = x
        for i, mod in enumerate(encoder):
            f = mod(f)
            if i == 1:
                pooled = f
        return f, pooled

    def forward(self, block, perms=None):
        B, N, C, T, H, W = block.shape
        assert N == 3
        s, e = self.n_series, self.series_dim
        x1, x2, aug = (block[:, i].contiguous() for i in range(3))
        fq, pooled_q = self._encode(self.encoder_q, None, x1)
        q = F.normalize(fq, dim=1).reshape(B, self.dim)
        series_q = F.normalize(self.series_proj_head_q(pooled_q).reshape(B, s, e), dim=2)
        train = q.requires_grad
        with torch.no_grad():
            if train:
                O.momentum_update(self.encoder_q.parameters(), self.encoder_k.parameters(), self.m)
                O.momentum_update(self.series_proj_head_q.parameters(), self.series_proj_head_k.parameters(), self.m)
            if self.distributed:
                x2, unshuf = self._shuffle_ddp(x2)
            fk, pooled_k = self._encode(self.encoder_k, None, x2)
            k = F.normalize(fk, dim=1)
            series_k = F.normalize(self.series_proj_head_k(pooled_k).reshape(-1, s, e), dim=2).reshape(-1, s * e)
            if self.distributed:
                k = self._unshuffle_ddp(k, unshuf)
                series_k = self._unshuffle_ddp(series_k, unshuf)
        k = k.reshape(B, self.dim)
        ret = _pack("clip_", *O.moco_infonce(q, k, self.queue, self.T))
        series_k = series_k.reshape(B, s, e)
        if self.with_tc:
            ret.update(_pack("tc_", *O.moco_tc(series_q, series_k, self.series_queue, self.aligned_T)))
        if train:
            keys, skeys = k, series_k.reshape(B, s * e)
            if self.distributed:
                keys, skeys = _all_gather_cat(keys), _all_gather_cat(skeys)
            ptr = int(self.queue_ptr)
            O.enqueue(self.queue, ptr, keys)
            self.queue_ptr[0] = O.enqueue(self.series_queue, ptr, skeys)
        if perms is None:
            perms = O.draw_segment_perms(B, s)
        dual = torch.cat([aug, O.shuffle_segments(aug, perms)], dim=0)
        g = dual
        for mod in list(self.encoder_q)[:2]:
            g = mod(g)
        dual_series = F.normalize(self.series_proj_head_q(g).reshape(2 * B, s, e), dim=2)
        aug_series, shuf = dual_series[:B], O.calibrate_segments(dual_series[B:], perms)
        for base, prefix in ((series_q, "unaug_ranking_"), (aug_series, "aug_ranking_")):
            pairs = torch.stack([base, shuf], dim=2)
            ret.update(_pack(prefix, *O.rank_loss(pairs, 0.05, 0.5, clip_max=None), margin=True))
        return ret


class LinearClassifier(nn.Module):
    """backbone -> global average pool -> [L2 norm] -> [BN1d] -> [dropout] -> Linear
    (model/classifier.py:9-84). Returns (logit, pooled feature)."""

    def __init__(self, num_class=101, network="resnet50", dropout=0.5, use_dropout=True, use_l2_norm=False,
                 use_final_bn=False, nonlinear=False, proj_dim=128):
        super().__init__()
        self.network, self.num_class = network, num_class
        self.use_l2_norm, self.use_final_bn = use_l2_norm, use_final_bn
        self.backbone, self.param = select_backbone(network)
        fs = self.param["feature_size"]
        if use_final_bn:
            self.final_bn = nn.BatchNorm1d(fs)
            self.final_bn.weight.data.fill_(1)
            self.final_bn.bias.data.zero_()
        if use_dropout:
            self.final_fc = nn.Sequential(nn.Dropout(dropout), nn.Linear(fs, num_class))
        elif nonlinear:
            self.final_fc = nn.Sequential(nn.Linear(fs, proj_dim), nn.ReLU(), nn.Linear(proj_dim, num_class))
        else:
            self.final_fc = nn.Sequential(nn.Linear(fs, num_class))
        for name, p in self.final_fc.named_parameters():
            if "bias" in name:
                nn.init.constant_(p, 0.0)
            elif "weight" in name:
                nn.init.normal_(p, mean=0.0, std=0.01)

    def forward(self, block):
        B = block.shape[0]
        feat = F.adaptive_avg_pool3d(self.backbone(block), (1, 1, 1)).reshape(B, self.param["feature_size"])
        if self.use_l2_norm:
            feat = F.normalize(feat, p=2, dim=1)
        logit = self.final_fc(self.final_bn(feat) if self.use_final_bn else feat)
        return logit, feat
