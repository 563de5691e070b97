"""Drop-in model wrappers: same class names, constructor signatures (positional order matters —
pretrain.py:61-77 passes positionally), ``state_dict`` keys and returned dict keys as the reference's
model/simclr.py, model/moco.py and model/classifier.py, running on the sm_100a kernels.

``forward(block)`` takes the reference's fp32 block (B, n_views, C, T, H, W) on a CUDA device and
returns ``{'<p>logits', '<p>labels', '<p>contrast_loss'}`` for p in clip_, tc_, aug_ranking_margin_,
unaug_ranking_margin_ (key spelling is load-bearing: pretrain.py:413-445). Losses are autograd
scalars; ``loss.backward()`` fills ``.grad`` of the fp32 parameters exactly like the reference.

The reference's forward calls a method that does not exist (``calc_contrast_loss``);
both spellings are provided here.
"""
import numpy as np
import torch
import torch.nn as nn

from . import engine as E
from . import objectives as O
from .backbones import select_backbone


def _proj_head(cin, cout):
    return [nn.Conv3d(cin, cin, kernel_size=1, bias=True), nn.ReLU(), nn.Conv3d(cin, cout, kernel_size=1, bias=True)]


def _head(mods, x):
    """[Conv3d 1x1x1, ReLU, Conv3d 1x1x1] on pooled (N, C) features."""
    return O.linear(O.linear(x, mods[0], relu=True), mods[2])


def _check_input(block):
    """Accept the reference's fp32 block or engine.RawClips (fused Normalize + transpose ingest)."""
    if isinstance(block, E.RawClips):
        if not block.device.type == "cuda":
            raise E._lib.DualVarNativeError("dualvar_b200 models run on a B200 only (no CPU fallback)")
        return block, block.block_shape
    if not block.is_cuda:
        raise E._lib.DualVarNativeError("dualvar_b200 models run on a B200 only (no CPU fallback)")
    block = block.contiguous().float()
    return block, tuple(block.shape)


def _draw_perms(B, n_series, device):
    """One np.random.permutation per sample from the global NumPy stream, in batch order
    (model/simclr.py:379-381, model/moco.py:544-546)."""
    perms = np.array([np.random.permutation(n_series) for _ in range(B)], dtype=np.int32)
    return torch.from_numpy(perms).to(device, non_blocking=True)


class SimCLR_Naked(nn.Module):
    """Two-view SimCLR (model/simclr.py:19-127)."""

    def __init__(self, network='s3d', dim=128, T=0.07, distributed=True, nonlinear=True):
        super().__init__()
        self.dim, self.distributed, self.T, self.nonlinear = dim, distributed, T, nonlinear
        backbone, self.param = select_backbone(network)
        feature_size = self.param['feature_size']
        self.encoder_q = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_q.extend(_proj_head(feature_size, dim))
        self.criterion = nn.CrossEntropyLoss()

    def calc_contrast_loss(self, features, n_views=2, prefix='clip_'):
        assert features.dim() == 3 and features.shape[1] == n_views == 2, features.shape
        ret, self.last_hits = O.nt_xent(features, self.T, self.distributed, prefix)
        return ret

    def forward(self, block):
        block, shape = _check_input(block)
        B, n_views = shape[:2]
        assert n_views == 2
        pooled = self.encoder_q[0].encode(block, pooled=True)
        f = _head(self.encoder_q[2:], pooled) if self.nonlinear else pooled
        f = O.l2norm(f).view(B, n_views, -1)
        return self.calc_contrast_loss(f, n_views, 'clip_')


class SimCLR_TimeSeriesV4(nn.Module):
    """SimCLR + DualVar (model/simclr.py:130-400)."""

    def __init__(self, network='s3d', dim=128, T=0.07, distributed=True, nonlinear=True, n_series=2, series_dim=64,
                 series_T=0.07, aligned_T=0.07, mode="clip-sr-tc", args=None):
        super().__init__()
        self.cnt = 0
        self.args = args
        self.dim, self.distributed, self.T, self.nonlinear = dim, distributed, T, nonlinear
        self.n_series, self.series_dim = n_series, series_dim
        self.series_T, self.aligned_T, self.mode = series_T, aligned_T, mode
        self.with_clip = 'clip' in mode
        self.with_sr = 'sr' in mode
        self.with_tc = 'tc' in mode
        backbone, self.param = select_backbone(network)
        feature_size = self.param['feature_size']
        self.encoder_q = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear and self.with_clip:
            self.encoder_q.extend(_proj_head(feature_size, dim))
        self.criterion = nn.CrossEntropyLoss()
        self.series_proj_head = nn.Sequential(*_proj_head(feature_size, series_dim * self.n_series))
        self.last_hits = {}

    # ---- objectives (same names as the reference methods) ----
    def calc_clip_contrast_loss(self, features, n_views=2, prefix='clip_'):
        assert features.dim() == 3 and features.shape[1] == n_views == 2, features.shape
        ret, self.last_hits[prefix] = O.nt_xent(features, self.T, self.distributed, prefix)
        return ret

    calc_contrast_loss = calc_clip_contrast_loss

    def calc_tc_contrast_loss(self, features, prefix="tc_"):
        B, n_views, n_series, dim = features.shape
        assert n_series == self.n_series and dim == self.series_dim
        ret, self.last_hits[prefix] = O.tc_loss(features, self.aligned_T, self.distributed, prefix)
        return ret

    def calc_ranking_loss(self, features, n_views=2, prefix='ranking_', weight=1.):
        """features (B, n_series, 2, series_dim) as in the reference (segment, view)."""
        assert features.dim() == 4 and features.shape[1] == self.n_series and features.shape[2] == n_views == 2
        a, b = features[:, :, 0].contiguous(), features[:, :, 1].contiguous()
        ret, self.last_hits[prefix] = O.rank_loss(a, b, self.args.shufflerank_theta, weight, 5.0, prefix)
        return ret

    def forward(self, block):
        block, shape = _check_input(block)
        B, n_views_in, C, T, H, W = shape
        assert n_views_in == 3
        s, e = self.n_series, self.series_dim
        backbone = self.encoder_q[0]
        dev = block.device
        # pass 1: all 3B clips in (b, view) order, one batch -> BN statistics over 3B (model/simclr.py:352-357)
        pooled = backbone.encode(block, pooled=True)                                    # (3B, fs)
        ret = dict()
        if self.with_clip:
            f = _head(self.encoder_q[2:], pooled) if len(self.encoder_q) > 2 else pooled
            f = O.l2norm(f).view(B, 3, -1)[:, :2]
            ret.update(self.calc_clip_contrast_loss(f, 2))
        series = O.l2norm(_head(self.series_proj_head, pooled).view(B * 3 * s, e)).view(B, 3, s, e)
        if self.with_tc:
            ret.update(self.calc_tc_contrast_loss(series[:, :2]))
        if self.with_sr:
            # pass 2: view 2 with its T/s-frame segments permuted per sample; the permutation is folded
            # into the ingest kernel's addressing (model/simclr.py:378-387)
            perm = _draw_perms(B, s, dev)
            pooled_s = backbone.encode(block, pooled=True, first_view=2, n_views=1, perm=perm, n_series=s)  # (B, fs)
            shuf = _head(self.series_proj_head, pooled_s).view(B, s, e)
            shuf = O.l2norm(O.PermuteSegmentsFn.apply(shuf, perm))
            theta = self.args.shufflerank_theta
            for view, prefix in ((0, 'aug_ranking_'), (2, 'unaug_ranking_')):
                r, self.last_hits[prefix] = O.rank_loss(series[:, view], shuf, theta, 0.5, 5.0, prefix)
                ret.update(r)
        return ret


class LinearClassifier(nn.Module):
    """backbone -> global average pool -> [L2 norm] -> [BN1d] -> [dropout] -> Linear
    (model/classifier.py:9-84); returns (logit, pooled feature). The encoder runs on the sm_100a
    kernels; the (B, 512) tail — final BN1d / dropout / fc — is a next-tier row
    and uses torch modules."""

    def __init__(self, num_class=101, network='resnet50', dropout=0.5, use_dropout=True, use_l2_norm=False,
                 use_final_bn=False, nonlinear=False, proj_dim=128):
        super().__init__()
        self.network, self.num_class, self.dropout = network, num_class, dropout
        self.use_dropout, self.use_l2_norm, self.use_final_bn = use_dropout, use_l2_norm, use_final_bn
        self.backbone, self.param = select_backbone(network)
        fs = self.param['feature_size']
        if use_final_bn:
            self.final_bn = nn.BatchNorm1d(fs)
            self.final_bn.weight.data.fill_(1)
            self.final_bn.bias.data.zero_()
        if use_dropout:
            self.final_fc = nn.Sequential(nn.Dropout(dropout), nn.Linear(fs, self.num_class))
        elif nonlinear:
            self.final_fc = nn.Sequential(nn.Linear(fs, proj_dim), nn.ReLU(), nn.Linear(proj_dim, self.num_class))
        else:
            self.final_fc = nn.Sequential(nn.Linear(fs, self.num_class))
        self._initialize_weights(self.final_fc)

    def forward(self, block):
        block, _ = _check_input(block)
        feat3d = self.backbone.encode(block, pooled=True)
        if self.use_l2_norm:
            feat3d = O.l2norm(feat3d)
        logit = self.final_fc(self.final_bn(feat3d) if self.use_final_bn else feat3d)
        return logit, feat3d

    def _initialize_weights(self, module):
        for name, param in module.named_parameters():
            if 'bias' in name:
                nn.init.constant_(param, 0.0)
            elif 'weight' in name:
                nn.init.normal_(param, mean=0.0, std=0.01)
