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


_perm_source = None     # graph_step.GraphedTrainStep: the permutations live in a static device buffer it refreshes


def _draw_perms(B, n_series, device):
    """One np.random.permutation per sample from the global NumPy stream, in batch order
    (model/simclr.py:379-381, model/moco.py:544-546)."""
    if _perm_source is not None:
        return _perm_source(B, n_series, device)
    perms = np.array([np.random.permutation(n_series) for _ in range(B)], dtype=np.int32)
    return torch.from_numpy(perms).to(device, non_blocking=True)


class SimCLR_Naked(nn.Module):
    """Two-view SimCLR (model/simclr.py:19-127)."""
    graph_safe = True      # a step can be captured as one CUDA graph (graph_step.GraphedTrainStep)

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
    graph_safe = True

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
        # pass 2 (below) is independent of pass 1's result: when enabled both are issued together on two streams
        # (the host-side permutation draw moves up; no other host RNG use lies between the two in the reference)
        perm = pooled_s = None
        if self.with_sr and E.PASS_STREAMS and not E.fp32_mode():
            perm = _draw_perms(B, s, dev)
            pooled, pooled_s = backbone.encode_pair(block, dict(), dict(first_view=2, n_views=1, perm=perm, n_series=s))
        else:
            pooled = backbone.encode(block, pooled=True)                                # (3B, fs)
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
            if pooled_s is None:
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
    kernels, the Linear layers of the (B, feature_size) tail on the sgemm kernel; BatchNorm1d and
    Dropout on that tiny tensor stay torch modules (Dropout must draw the reference's mask)."""
    graph_safe = True

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
        h = self.final_bn(feat3d) if self.use_final_bn else feat3d
        # final_fc: Linear layers on the sgemm kernel (bias / ReLU fused in its epilogue); Dropout stays torch's so the
        # mask comes from the same Philox stream as the reference's (model/classifier.py:39-52)
        mods = list(self.final_fc)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                fuse = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                h = O.linear(h, m, relu=fuse)
                i += 2 if fuse else 1
            else:
                h = m(h)
                i += 1
        return h, feat3d

    def _initialize_weights(self, module):
        for name, param in module.named_parameters():
            if 'bias' in name:
                nn.init.constant_(param, 0.0)
            elif 'weight' in name:
                nn.init.normal_(param, mean=0.0, std=0.01)


# ------------------------------------------------------------------------------------------- MoCo
@torch.no_grad()
def concat_all_gather(tensor):
    """all_gather without gradient (model/moco.py:14-25)."""
    import torch.distributed as dist
    out = torch.empty((dist.get_world_size() * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype,
                      device=tensor.device)
    dist.all_gather_into_tensor(out, tensor.contiguous())
    return out


class _MoCoBase(nn.Module):
    """Shared MoCo machinery: momentum update (one multi-tensor kernel), queue enqueue, shuffle-BN."""

    def _pairs(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def _momentum_table(self):
        """Device table (k ptr, q ptr, count) in 8192-element chunks; rebuilt if storage moves."""
        import ctypes
        pairs = [(pq, pk) for pq, pk in self._pairs()]
        sig = tuple((pq.data_ptr(), pk.data_ptr(), pq.numel()) for pq, pk in pairs)
        if getattr(self, "_mom_sig", None) != sig:
            rows = []
            for pq, pk in pairs:
                n = pq.numel()
                for off in range(0, n, 8192):
                    rows.append((pk.data_ptr() + 4 * off, pq.data_ptr() + 4 * off, min(8192, n - off)))
            self._mom_table = torch.tensor(rows, dtype=torch.int64).to(pairs[0][0].device)
            self._mom_sig = sig
        return self._mom_table

    @torch.no_grad()
    def _momentum_update_key_encoder(self):
        """theta_k = m*theta_k + (1-m)*theta_q (model/moco.py:328-334) — one launch for all tensors."""
        import ctypes
        t = self._momentum_table()
        E.call("dv_moco_momentum_update", E.ptr(t), t.shape[0], ctypes.c_float(self.m), E.stream_ptr())
        E.invalidate_weights(pk for _, pk in self._pairs())

    @property
    def graph_safe(self):
        """A step can be replayed as one CUDA graph (graph_step.GraphedTrainStep): the queue pointer never leaves the
        device. Not under shuffle-BN, whose permutation is drawn on the host every step (model/moco.py:371-381)."""
        return not self._distributed_on()

    def _advance_ptr(self, batch):
        """queue_ptr = (queue_ptr + batch) % K on the device (model/moco.py:352-353; the reference reads the pointer
        back with int(self.queue_ptr) every step - a device-to-host sync this path does not have)."""
        E.call("dv_moco_advance_ptr", E.ptr(self.queue_ptr), batch, self.K, E.stream_ptr())

    @torch.no_grad()
    def _enqueue(self, queue, keys):
        """queue[:, ptr:ptr+B] = keys^T with ptr read from self.queue_ptr on the device (model/moco.py:343-351)."""
        B, d = keys.shape
        assert self.K % B == 0  # for simplicity (model/moco.py:347)
        E.call("dv_moco_enqueue_at", E.ptr(keys.contiguous()), E.ptr(queue), B, d, self.K, E.ptr(self.queue_ptr),
               E.stream_ptr())

    def _distributed_on(self):
        return O._dist_on(self.distributed)

    @torch.no_grad()
    def _shuffle_plan(self, n_local):
        """idx_shuffle from rank 0 (model/moco.py:371-374) turned into an all-to-all plan (shuffle_bn_plan): every rank
        draws a permutation (the reference's RNG consumption), rank 0's wins; one 8-byte-per-clip read-back gives the
        host the split sizes."""
        import torch.distributed as dist
        world = dist.get_world_size()
        idx = torch.randperm(n_local * world).cuda()
        dist.broadcast(idx, src=0)
        return shuffle_bn_plan(idx.cpu(), world, dist.get_rank(), n_local, idx.device)

    @torch.no_grad()
    def _unshuffle(self, x, rows):
        """Keys back in the original order (model/moco.py:385-402): gather every rank's keys (128 floats per clip) and
        take this rank's rows."""
        return concat_all_gather(x)[rows]

    def _key_input(self, backbone, block, view):
        """Ingest the key clips; with shuffle-BN every rank SENDS each of its bf16 ingested clips to the one rank whose
        BatchNorm group the permutation puts it in (one all-to-all: B clips leave and B clips arrive per rank, instead of
        gathering world * B clips and dropping all but B; model/moco.py:357-383)."""
        s2d = backbone.wants_s2d(block)
        if not self._distributed_on():
            return (lambda: E.ingest(block, first_view=view, n_views=1, s2d=s2d)), None
        B = (block.block_shape if isinstance(block, E.RawClips) else block.shape)[0]
        plan = self._shuffle_plan(B)

        def make():
            local = E.ingest(block, first_view=view, n_views=1, s2d=s2d)
            cd = E.clip_dim()      # fp32 mode: split planes [K][clips][...]
            t = E.input_tensor(local)
            if cd == 0:
                sel = exchange_clips(t, plan)
            else:
                sel = exchange_clips(t.transpose(0, 1).contiguous(), plan).transpose(0, 1).contiguous()
            return E.input_act(sel, local.C, local.s2d)
        return make, plan["unshuffle_rows"]


def shuffle_bn_plan(idx, world, rank, n_local, device=None):
    """All-to-all plan of MoCo's shuffle-BN for one rank.

    ``idx`` (host, int64, world * n_local) is the reference's ``idx_shuffle``: rank g's key encoder sees the clips
    ``idx.view(world, -1)[g]`` of the gathered batch (model/moco.py:376-381). Clip j lives on rank j // n_local, so
    rank r sends to rank g its clips {j % n_local : j in idx_g, j // n_local == r}. BatchNorm statistics do not depend
    on the order of the clips inside a rank's batch, so a rank keeps its share in ARRIVAL order (by source rank, then
    by position in idx_g) - no reordering copy - and the un-shuffle indices account for that order: the gathered keys
    come back exactly in the original order, as with the reference's ``idx_unshuffle`` (model/moco.py:377,400).
    Returns send_index (local clip indices in send order), send_counts / recv_counts per peer, and unshuffle_rows (rows of
    the all-gathered keys that are this rank's original clips)."""
    idx = idx.view(world, n_local)
    src = idx // n_local
    # arrival order on rank g: stable sort of idx_g by source rank
    order = torch.argsort(src, dim=1, stable=True)
    eff = torch.gather(idx, 1, order)                       # eff[g][q] = original global clip in slot q of rank g
    send_index = torch.cat([(eff[g][eff[g] // n_local == rank]) % n_local for g in range(world)])
    send_counts = [int((eff[g] // n_local == rank).sum()) for g in range(world)]
    recv_counts = [int((eff[rank] // n_local == r).sum()) for r in range(world)]
    unshuffle = torch.argsort(eff.reshape(-1))              # position of original clip j among the gathered keys
    rows = unshuffle.view(world, n_local)[rank]
    if device is not None:
        send_index, rows = send_index.to(device), rows.to(device)
    return {"send_index": send_index, "send_counts": send_counts, "recv_counts": recv_counts, "unshuffle_rows": rows,
            "effective": eff}


@torch.no_grad()
def exchange_clips(t, plan):
    """This rank's shuffled share of the clips ``t`` [n_local, ...] (see shuffle_bn_plan): one all_to_all_single."""
    import torch.distributed as dist
    send = t.index_select(0, plan["send_index"]).contiguous()
    out = torch.empty_like(t)
    dist.all_to_all_single(out, send, output_split_sizes=plan["recv_counts"], input_split_sizes=plan["send_counts"])
    return out


class MoCo_Naked(_MoCoBase):
    """Two-view MoCo (model/moco.py:28-239)."""

    def __init__(self, network='s3d', dim=128, K=2048, m=0.999, T=0.07, distributed=True, nonlinear=True):
        super().__init__()
        self.dim, self.K, self.m, self.T = dim, K, m, T
        self.distributed, self.nonlinear = distributed, nonlinear
        backbone, self.param = select_backbone(network)
        feature_size = self.param['feature_size']
        self.encoder_q = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_q.extend(_proj_head(feature_size, dim))
        backbone, _ = select_backbone(network)
        self.encoder_k = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_k.extend(_proj_head(feature_size, dim))
        for param_q, param_k in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            param_k.data.copy_(param_q.data)
            param_k.requires_grad = False
        self.register_buffer("queue", torch.randn(dim, K))
        self.queue = nn.functional.normalize(self.queue, dim=0)
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
        self.criterion = nn.CrossEntropyLoss()

    def _pairs(self):
        return zip(self.encoder_q.parameters(), self.encoder_k.parameters())

    def forward(self, block):
        block, shape = _check_input(block)
        B, N = shape[:2]
        assert N == 2
        bq, bk = self.encoder_q[0], self.encoder_k[0]
        pooled_q = bq.encode(block, pooled=True, first_view=0, n_views=1)
        q = O.l2norm(_head(self.encoder_q[2:], pooled_q) if self.nonlinear else pooled_q)
        in_train_mode = q.requires_grad
        with torch.no_grad():
            if in_train_mode:
                self._momentum_update_key_encoder()
            make, unshuf = self._key_input(bk, block, 1)
            pooled_k = bk.encode(None, pooled=True, make_input=make)
            k = O.l2norm(_head(self.encoder_k[2:], pooled_k) if self.nonlinear else pooled_k)
            if unshuf is not None:
                k = self._unshuffle(k, unshuf)
        ret, self.last_hits = O.queue_contrast(q, k, self.queue, self.T, 'clip_')
        if in_train_mode:
            keys = concat_all_gather(k) if self._distributed_on() else k
            self._enqueue(self.queue, keys)
            self._advance_ptr(keys.shape[0])
        return ret


class MoCo_TimeSeriesV4(_MoCoBase):
    """MoCo + DualVar (model/moco.py:242-573)."""

    def __init__(self, network='s3d', dim=128, K=2048, m=0.999, T=0.07, distributed=True, nonlinear=True,
                 n_series=2, series_dim=64, series_T=0.07, aligned_T=0.07, mode="clip-sr-tc", args=None):
        super().__init__()
        self.dim, self.K, self.m, self.T = dim, K, m, T
        self.distributed, self.nonlinear = distributed, nonlinear
        self.n_series, self.series_dim, self.mode = n_series, series_dim, mode
        self.series_T, self.aligned_T = series_T, aligned_T
        self.with_clip = 'clip' in mode
        self.with_sr = 'sr' in mode
        self.with_tc = 'tc' in mode
        backbone, self.param = select_backbone(network)
        feature_size = self.param['feature_size']
        self.encoder_q = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_q.extend(_proj_head(feature_size, dim))
        self.series_proj_head_q = nn.Sequential(*_proj_head(feature_size, series_dim * n_series))
        backbone, _ = select_backbone(network)
        self.encoder_k = nn.ModuleList([backbone, nn.AdaptiveAvgPool3d((1, 1, 1))])
        if nonlinear:
            self.encoder_k.extend(_proj_head(feature_size, dim))
        self.series_proj_head_k = nn.Sequential(*_proj_head(feature_size, series_dim * n_series))
        for param_q, param_k in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            param_k.data.copy_(param_q.data)
            param_k.requires_grad = False
        for param_q, param_k in zip(self.series_proj_head_q.parameters(), self.series_proj_head_k.parameters()):
            param_k.data.copy_(param_q.data)
            param_k.requires_grad = False
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
        self.register_buffer("queue", torch.randn(dim, K))
        self.queue = nn.functional.normalize(self.queue, dim=0)
        self.register_buffer("series_queue", torch.randn(series_dim * n_series, K))
        self.series_queue = nn.functional.normalize(
            self.series_queue.view(n_series, series_dim, K), dim=1).view(n_series * series_dim, K)
        self.criterion = nn.CrossEntropyLoss()
        self.last_hits = {}

    def _pairs(self):
        import itertools
        return itertools.chain(zip(self.encoder_q.parameters(), self.encoder_k.parameters()),
                               zip(self.series_proj_head_q.parameters(), self.series_proj_head_k.parameters()))

    # objectives under the reference's method names
    def calc_clip_contrast_loss(self, q, k, queue, prefix='clip_'):
        ret, self.last_hits[prefix] = O.queue_contrast(q, k, queue, self.T, prefix)
        return ret

    calc_contrast_loss = calc_clip_contrast_loss

    def calc_tc_contrast_loss(self, q, k, queue, prefix="tc_"):
        """mean over s x s segment pairs == dot product of segment means (model/moco.py:404-424)."""
        B, s, e = q.shape
        assert s == self.n_series and e == self.series_dim
        qm = O.SegmentMeanFn.apply(q)
        with torch.no_grad():
            km = O.SegmentMeanFn.apply(k.reshape(B, s, e))
            queue_m = O.SegmentMeanFn.apply(queue.detach().view(1, s, e * self.K)).view(e, self.K)
        ret, self.last_hits[prefix] = O.queue_contrast(qm, km, queue_m, self.aligned_T, prefix)
        return ret

    def calc_ranking_loss(self, features, n_views=2, prefix='ranking_', weight=1.):
        a, b = features[:, :, 0].contiguous(), features[:, :, 1].contiguous()
        ret, self.last_hits[prefix] = O.rank_loss(a, b, 0.05, weight, None, prefix)
        return ret

    def forward(self, block):
        block, shape = _check_input(block)
        B, N, C, T, H, W = shape
        assert N == 3
        s, e = self.n_series, self.series_dim
        bq, bk = self.encoder_q[0], self.encoder_k[0]
        dev = block.device
        ret = {}
        pooled_q = bq.encode(block, pooled=True, first_view=0, n_views=1)
        q = O.l2norm(_head(self.encoder_q[2:], pooled_q) if self.nonlinear else pooled_q)
        series_q = O.l2norm(_head(self.series_proj_head_q, pooled_q).view(B * s, e)).view(B, s, e)
        in_train_mode = q.requires_grad
        with torch.no_grad():
            if in_train_mode:
                self._momentum_update_key_encoder()
            make, unshuf = self._key_input(bk, block, 1)
            pooled_k = bk.encode(None, pooled=True, make_input=make)
            k = O.l2norm(_head(self.encoder_k[2:], pooled_k) if self.nonlinear else pooled_k)
            series_k = O.l2norm(_head(self.series_proj_head_k, pooled_k).view(B * s, e)).view(B, s * e)
            if unshuf is not None:
                k = self._unshuffle(k, unshuf)
                series_k = self._unshuffle(series_k, unshuf)
        ret.update(self.calc_contrast_loss(q, k, self.queue, 'clip_'))
        if self.with_tc:
            ret.update(self.calc_tc_contrast_loss(series_q, series_k.view(B, s, e), self.series_queue, 'tc_'))
        if in_train_mode:
            keys, skeys = k, series_k
            if self._distributed_on():
                keys, skeys = concat_all_gather(keys), concat_all_gather(skeys)
            self._enqueue(self.queue, keys)
            self._enqueue(self.series_queue, skeys)
            self._advance_ptr(keys.shape[0])
        # view 2 twice in one batch: as is, and with its segments shuffled (model/moco.py:543-557)
        perm = _draw_perms(B, s, dev)
        s2d = bq.wants_s2d(block)

        def make_dual():
            buf = E.ingest_buffer(block, 2 * B, s2d, dev)
            a = E.ingest(block, first_view=2, n_views=1, s2d=s2d, out=E.clip_range(buf, 0, B))
            E.ingest(block, first_view=2, n_views=1, perm=perm, n_series=s, s2d=s2d, out=E.clip_range(buf, B, 2 * B))
            return E.input_act(buf, a.C, (2 * B,) + a.s2d[1:] if a.s2d else None)

        pooled_d = bq.encode(None, pooled=True, make_input=make_dual)
        series_d = O.l2norm(_head(self.series_proj_head_q, pooled_d).view(2 * B * s, e)).view(2 * B, s, e)
        aug_series = series_d[:B]
        shuf = O.PermuteSegmentsFn.apply(series_d[B:], perm)
        for base, prefix in ((series_q, 'unaug_ranking_'), (aug_series, 'aug_ranking_')):
            r, self.last_hits[prefix] = O.rank_loss(base, shuf, 0.05, 0.5, None, prefix)
            ret.update(r)
        return ret
