"""Forward/backward executor for the clip encoders on bf16 NDHWC activations.

The reference runs its backbones op by op through autograd (conv -> BN -> ReLU ..., each an ATen /
cuDNN call that reads and writes a full fp32 NCDHW tensor). Here a backbone forward is ONE
``torch.autograd.Function``: the host modules (dualvar_b200/backbones.py) describe their graph with the
primitives below, every primitive launches kernels through the C ABI and records a closure on a tape;
the Function's backward replays the tape in reverse. torch only allocates memory and carries the
fp32 parameters/gradients at the boundary.

Primitive <-> reference op:
  conv_stats   nn.Conv3d forward + the batch statistics of the nn.BatchNorm3d that follows
  activate     BatchNorm3d normalise/affine (+ second BN branch) (+ residual) (+ ReLU)
  max_pool     nn.MaxPool3d
  global_pool  nn.AdaptiveAvgPool3d((1,1,1))
"""
import ctypes
import os
import weakref

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib, comm
from ._lib import make_geom, pad8, ptr, stream_ptr

call = _lib.call

# ----------------------------------------------------------------------------- precision mode
# "bf16" (default, the throughput mode): bf16 NDHWC activations, bf16 x bf16 -> fp32 tensor-core convolutions.
# "fp32" (the 1e-4 parity mode of the north star): fp32 NDHWC activations and gradients; every conv operand is also
# kept as F32_PLANES bf16 split planes (x = p0 + p1 + p2, csrc/fp32_mode.cu) and a convolution is the sum of the
# plane products with i + j < F32_PLANES, each one launch of the same tcgen05 kernel adding its fp32 accumulator tile
# to the fp32 output. 3 planes carry 24 mantissa bits (6 launches per conv), 2 planes 16 bits (3 launches).
# Covers every select_backbone network (r21d / r3d / c3d / s3d / s3dg / r2d3d18).
PRECISION = os.environ.get("DV_PRECISION", "bf16")
F32_PLANES = int(os.environ.get("DV_FP32_PLANES", "3"))


def set_precision(mode, planes=None):
    """Select the arithmetic of the encoder passes built from now on: "bf16" or "fp32" (see above)."""
    global PRECISION, F32_PLANES
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    if planes is not None:
        if planes not in (1, 2, 3):
            raise ValueError("fp32 mode uses 1, 2 or 3 split planes")
        F32_PLANES = int(planes)
    PRECISION = mode


def fp32_mode():
    return PRECISION == "fp32"


def _terms():
    """Plane products (operand plane i, weight plane j) of one convolution, smallest contributions first."""
    K = F32_PLANES
    return sorted(((i, j) for i in range(K) for j in range(K - i)), key=lambda t: -(t[0] + t[1]))


class Act:
    """An NDHWC activation [N,T,H,W,Cp] with C logical channels and an optional gradient: bf16 ``data`` in the bf16
    mode; in the fp32 mode fp32 ``data`` (None for the ingested clips) plus its bf16 split ``planes`` [K,N,T,H,W,Cp]."""

    __slots__ = ("data", "C", "grad", "grad2", "needs_grad", "s2d", "bnred", "fused", "planes", "lazy")

    def __init__(self, data, C, needs_grad=True, s2d=None, planes=None, lazy=None):
        self.data = data
        self.planes = planes
        self.lazy = lazy        # (RawBN, relu): z = relu?(BN(y)) exists only inside its consumer conv (never in HBM)
        self.C = C
        self.grad = None
        self.grad2 = None       # second pending contribution (summed lazily by the consumer)
        self.needs_grad = needs_grad
        self.s2d = s2d          # (N, T, H, W) of the original frames when data is the space-to-depth stem layout
        self.bnred = None       # (RawBN, relu) when this is a plain relu?(BN(y)): lets a consumer's dgrad fuse the BN-backward reduce
        self.fused = None       # (sums, dx): BN-backward sums a dgrad epilogue already produced for gradient dx

    @property
    def shape5(self):
        if self.lazy is not None:
            return tuple(self.lazy[0].y.shape)
        return tuple(self.data.shape) if self.data is not None else tuple(self.planes.shape[1:])

    @property
    def stored(self):
        """The tensor in HBM that stands for this activation: z itself, or the raw y of a lazy one."""
        return self.lazy[0].y if self.lazy is not None else self.data

    @property
    def rows(self):
        s = self.shape5
        return s[0] * s[1] * s[2] * s[3]

    @property
    def Cp(self):
        return self.shape5[4]

    @property
    def device(self):
        if self.lazy is not None:
            return self.lazy[0].y.device
        return self.data.device if self.data is not None else self.planes.device


class RawBN:
    """Un-normalised conv output + everything BN needs, between conv_stats() and activate()."""

    __slots__ = ("y", "ss", "saved", "geom", "x", "conv", "bn", "packed", "count", "sync", "stem")


# Direct parameter gradients (graph_step.GraphedTrainStep, single process): inside ``direct_param_grads(params)`` the
# weight-gradient unpack and the BatchNorm-backward finalisation ADD their results straight into the parameters' existing
# (zeroed) ``.grad`` tensors and the backbone pass returns None for them - autograd then neither sums the two passes of a
# step (one add kernel per parameter) nor adds that sum into ``.grad`` (another one): ~160 tiny kernels per R(2+1)D step,
# ~600 per S3D-G step. Outside the context gradients go back through autograd as usual (DDP hooks, fit()'s eager loop).
_direct_ids = None


class direct_param_grads:
    def __init__(self, params):
        self.ids = {id(p) for p in params if p.requires_grad and p.grad is not None and p.grad.is_contiguous()}

    def __enter__(self):
        global _direct_ids
        self.prev = _direct_ids
        _direct_ids = self.ids
        return self

    def __exit__(self, *exc):
        global _direct_ids
        _direct_ids = self.prev
        return False


class Context:
    """Per-forward state: tape of backward closures, parameter-gradient sink, mode flags."""

    def __init__(self, training, record=True):
        self.training = training
        self.record = record
        self.tape = []
        self.param_grads = {}      # id(param) -> fp32 tensor shaped like the parameter
        self.overrides = {}        # (id(concat Act), channel offset) -> dense gradient replacing the slice's
        self.side_used = False     # weight gradients are in flight on the side stream
        self.defer_running = None  # list: running-statistic updates of a concurrent pass, applied in order at the join
        self.nbt_pending = []      # num_batches_tracked buffers of the BatchNorm layers this pass has updated
        self._arena = None         # zeroed fp64 scratch the per-layer statistic buffers are carved from
        self._arena_used = 0
        self.rpass = None          # parallel._Pass during backward: parameter gradients go into its flat buckets
        self.direct = set()        # ids of parameters whose gradient was added straight into .grad (direct_param_grads)

    def direct_target(self, p):
        """p.grad if this backward may add p's gradient into it in place (direct_param_grads), else None."""
        if _direct_ids is None or self.rpass is not None or id(p) not in _direct_ids or id(p) in self.param_grads:
            return None
        return p.grad

    def zeros64(self, n, device):
        """A zero-initialised float64 vector of n elements: one memset per ~64 K elements instead of one tiny
        fill kernel per BatchNorm layer and pass (statistics and backward sums are 2*Cp doubles each)."""
        n_al = (n + 31) // 32 * 32
        if self._arena is None or self._arena.device != device or self._arena_used + n_al > self._arena.numel():
            self._arena = torch.zeros(max(65536, n_al), dtype=torch.float64, device=device)
            self._arena_used = 0
        out = self._arena[self._arena_used:self._arena_used + n]
        self._arena_used += n_al
        return out

    def add_param_grad(self, p, g):
        k = id(p)
        if k in self.param_grads:
            if self.rpass is not None and self.rpass.in_flight(p):
                self.rpass.extra.append((p, g))        # its bucket is already being all-reduced
            else:
                self.param_grads[k].add_(g)
        else:
            self.param_grads[k] = g

    def grad_out(self, p):
        """Destination for p's gradient: its slice of the data-parallel reducer's flat bucket buffer during a reduced
        backward (the all-reduce then needs no gather copy), a fresh tensor otherwise."""
        v = self.rpass.view(p) if self.rpass is not None else None
        return v if v is not None else torch.empty_like(p)

    def grad_ready(self, p, stream=None):
        """p's gradient is complete once the work queued so far on ``stream`` has run: its bucket may be reduced."""
        if self.rpass is not None:
            self.rpass.ready(p, stream)


# ----------------------------------------------------------------------------- helpers
_weight_cache = {}
# Fused dgrad + BN-backward reduce (dv_conv3d_dgrad_bnred_bf16): 1 = where it pays (default), 0 = never, 2 = always.
# The fused column pass costs the dgrad epilogue ~1500 cycles per 64-channel chunk of dx (two chunks in flight);
# that hides behind the MMA main loop of the K-heavy (spatial / full 3-D) dgrads but not behind the temporal ones,
# where the separate HBM pass is cheaper (profiles/r01_fused_bn_reduce.txt).
FUSE_BN_REDUCE = int(os.environ.get("DV_FUSE_BN_REDUCE", "1"))


def _fuse_reduce_pays(g):
    if FUSE_BN_REDUCE != 1:
        return FUSE_BN_REDUCE == 2
    mma_cycles = g.taps * (g.Cout_p / 16.0) * max(g.Cin_p / 2.0, 53.0)     # per 128-position tile (tests/diag/mma_rate.py)
    epi_cycles = -(-g.Cin_p // 64) * 1500 / 2.0
    return mma_cycles >= 1.5 * epi_cycles


def _cache_get(key, w, ver):
    """Cached entry for parameter ``w`` under ``key``, or None. An entry belongs to one live parameter object: it
    holds a weak reference to it and is dropped when the parameter is freed, so a new model whose parameters reuse
    the ids / addresses of a deleted one (CPython and the caching allocator both recycle) never sees its packs."""
    hit = _weight_cache.get(key)
    if hit is not None and hit[0]() is w and hit[1] == ver:
        return hit[2]
    return None


def _cache_put(key, w, ver, payload):
    if key not in _weight_cache or _weight_cache[key][0]() is not w:
        weakref.finalize(w, _cache_drop, key, id(w))
    _weight_cache[key] = (weakref.ref(w), ver, payload)


def _cache_drop(key, wid):
    hit = _weight_cache.get(key)
    if hit is not None and hit[0]() is None:      # not already replaced by a live parameter with the same id
        _weight_cache.pop(key, None)


def packed_weights(conv):
    """bf16 packed copies of a conv weight, refreshed when the fp32 parameter changes
    (optimizer.step() bumps Tensor._version)."""
    w = conv.weight
    key = id(w)
    ver = (w._version, w.data_ptr())
    hit = _cache_get(key, w, ver)
    if hit is not None:
        _after_pack(hit[2])
        return hit[0], hit[1]
    Cout, Cin, kt, kh, kw = w.shape
    g = make_geom(1, kt, kh, kw, Cin, Cout, (kt, kh, kw), (1, 1, 1), (0, 0, 0))
    wf = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.bfloat16, device=w.device)
    wt = torch.empty((g.Cin_p, g.taps, g.Cout_p), dtype=torch.bfloat16, device=w.device)
    call("dv_pack_conv_weight", ptr(w.detach()), ptr(wf), ptr(wt), ctypes.byref(g), stream_ptr())
    _cache_put(key, w, ver, (wf, wt, _pack_mark()))
    return wf, wt


def _pack_mark():
    """(stream, event) of a pack kernel: a consumer on another stream (second backbone pass) orders itself after it."""
    st = torch.cuda.current_stream()
    return st, st.record_event()


def _after_pack(mark):
    st, ev = mark
    cur = torch.cuda.current_stream()
    if cur != st:
        cur.wait_event(ev)


def invalidate_weights(params):
    """Drop cached packed copies of parameters that were modified through raw pointers (kernels do not
    bump Tensor._version)."""
    for p in params:
        for key in (id(p), ("stem", id(p)), ("f32", id(p)), ("f32stem", id(p)), ("f32all", id(p))):
            hit = _weight_cache.get(key)
            if hit is not None:      # keep the weak reference (and its finalizer), drop version and payload
                _weight_cache[key] = (hit[0], None, None)


def packed_stem_weights(conv, g):
    """bf16 [Cout_p][kt*4][64] stem weights for the space-to-depth formulation."""
    w = conv.weight
    key = ("stem", id(w))
    ver = (w._version, w.data_ptr())
    hit = _cache_get(key, w, ver)
    if hit is not None:
        _after_pack(hit[1])
        return hit[0]
    ws = torch.empty((g.Cout_p, g.kt * 4, 64), dtype=torch.bfloat16, device=w.device)
    call("dv_pack_stem_weight", ptr(w.detach()), ptr(ws), ctypes.byref(g), stream_ptr())
    _cache_put(key, w, ver, (ws, _pack_mark()))
    return ws


def _split_weight(w, K):
    """fp32 [K][numel]: the bf16-representable parts of the weight (the pack kernels then round them exactly)."""
    planes = torch.empty((K, w.numel()), dtype=torch.float32, device=w.device)
    call("dv_f32_split_planes", ptr(w.detach().contiguous()), ptr(planes), w.numel(), K, stream_ptr())
    return planes


def packed_weight_planes(conv):
    """fp32 mode: [(wf_k, wt_k)] bf16 packed copies of the K split planes of a conv weight."""
    w = conv.weight
    K = F32_PLANES
    key = ("f32", id(w))
    ver = (w._version, w.data_ptr(), K)
    hit = _cache_get(key, w, ver)
    if hit is not None:
        _after_pack(hit[1])
        return hit[0]
    Cout, Cin, kt, kh, kw = w.shape
    g = make_geom(1, kt, kh, kw, Cin, Cout, (kt, kh, kw), (1, 1, 1), (0, 0, 0))
    planes = _split_weight(w, K)
    out = []
    for k in range(K):
        wf = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.bfloat16, device=w.device)
        wt = torch.empty((g.Cin_p, g.taps, g.Cout_p), dtype=torch.bfloat16, device=w.device)
        call("dv_pack_conv_weight", ptr(planes[k]), ptr(wf), ptr(wt), ctypes.byref(g), stream_ptr())
        out.append((wf, wt))
    _cache_put(key, w, ver, (out, _pack_mark()))
    return out


# fp32 mode: all plane products of a convolution in ONE launch (dv_conv3d_fprop_f32planes / dv_conv3d_dgrad_f32planes):
# the products are extra taps of one TMEM accumulator and the fp32 output is stored once, instead of one launch per
# product with a read-modify-write of the output for all but the first. DV_F32_MERGE=0 restores the per-product launches.
F32_MERGE = os.environ.get("DV_F32_MERGE", "1") != "0"


def packed_weight_planes_all(conv):
    """fp32 mode: (wf_all [Cout_p][K*taps][Cin_p], wt_all [Cin_p][K*taps][Cout_p]) - the K packed planes side by side
    along the tap dimension (plane j in tap slots [j*taps, (j+1)*taps)), for the merged plane-product launches."""
    w = conv.weight
    key = ("f32all", id(w))
    ver = (w._version, w.data_ptr(), F32_PLANES)
    hit = _cache_get(key, w, ver)
    if hit is not None:
        _after_pack(hit[1])
        return hit[0]
    planes = packed_weight_planes(conv)
    out = (torch.cat([p_[0] for p_ in planes], 1).contiguous(), torch.cat([p_[1] for p_ in planes], 1).contiguous())
    _cache_put(key, w, ver, (out, _pack_mark()))
    return out


def packed_stem_weight_planes(conv, g):
    """fp32 mode: [(ws_k, None)] space-to-depth stem packs of the K split planes."""
    w = conv.weight
    K = F32_PLANES
    key = ("f32stem", id(w))
    ver = (w._version, w.data_ptr(), K)
    hit = _cache_get(key, w, ver)
    if hit is not None:
        _after_pack(hit[1])
        return hit[0]
    planes = _split_weight(w, K)
    out = []
    for k in range(K):
        ws = torch.empty((g.Cout_p, g.kt * 4, 64), dtype=torch.bfloat16, device=w.device)
        call("dv_pack_stem_weight", ptr(planes[k]), ptr(ws), ctypes.byref(g), stream_ptr())
        out.append((ws, None))
    _cache_put(key, w, ver, (out, _pack_mark()))
    return out


def stem_eligible(conv, H, W):
    """Stride-2 7x7 first conv on <=4 channels with even frame size -> space-to-depth stem kernels."""
    return (conv.weight.shape[1] <= 4 and tuple(conv.weight.shape[3:]) == (7, 7) and tuple(conv.stride) == (1, 2, 2)
            and tuple(conv.padding)[1:] == (3, 3) and H % 2 == 0 and W % 2 == 0)


def _bias_padded(conv, Cp):
    if conv.bias is None:
        return None
    b = torch.zeros(Cp, dtype=torch.float32, device=conv.bias.device)
    b[:conv.bias.shape[0]] = conv.bias.detach()
    return b


def _is_sync(bn):
    return isinstance(bn, nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized() \
        and dist.get_world_size() > 1


def _add_into(a, b):
    """a += b on the gradient's own element type (bf16 mode / fp32 mode)."""
    call("dv_f32_add" if a.dtype == torch.float32 else "dv_add_bf16", ptr(a), ptr(b), ptr(a), a.numel(), stream_ptr())


def _acc_grad(act, g):
    """Accumulate a gradient contribution. The second contribution is kept separate: the BN-backward
    kernels sum two streams on the fly, so the usual "main path + shortcut" meeting needs no add pass."""
    if act.grad is None:
        act.grad = g
    elif act.grad2 is None and g.shape == act.grad.shape:
        act.grad2 = g
    else:
        _materialize_grad(act)
        _add_into(act.grad, g)


def _materialize_grad(act):
    """Fold the pending second contribution into act.grad (consumers that read a single tensor)."""
    if act.grad2 is not None:
        _add_into(act.grad, act.grad2)
        act.grad2 = None
    return act.grad


# ----------------------------------------------------------------------------- primitives
class RawClips:
    """Un-normalised loader output (B, C, V*T, H, W) in [0,1] plus the Normalize constants: lets the
    models fuse pretrain.py:386-389 (Normalize + view + transpose + contiguous) into the ingest kernel
    instead of three fp32 passes. ``model(RawClips(...))`` == ``model(tr(x))`` of the reference loop."""

    def __init__(self, frames, n_views, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        assert frames.dim() == 5 and frames.shape[2] % n_views == 0
        # uint8 frames (decoded images) are converted on the GPU as ToTensor would on the host (x / 255)
        self.frames = frames.contiguous() if frames.dtype == torch.uint8 else frames.contiguous().float()
        self.n_views = n_views
        self.mean, self.std = tuple(mean), tuple(std)

    @property
    def device(self):
        return self.frames.device

    @property
    def block_shape(self):
        B, C, VT, H, W = self.frames.shape
        return (B, self.n_views, C, VT // self.n_views, H, W)


def ingest(src, first_view=0, n_views=None, perm=None, n_series=0, s2d=False, out=None):
    """Clips -> bf16 NDHWC Act with 8 channels. ``src`` is the reference block (B, V, C, T, H, W)
    fp32, a plain clip batch (B, C, T, H, W), or RawClips. Output clip order is (b, view) with views
    first_view .. first_view+n_views-1, matching block.view(-1, C, T, H, W) (model/simclr.py:352)."""
    mean = std = None
    if isinstance(src, RawClips):
        t = src.frames
        B, V, C, T, H, W = src.block_shape
        sb, sv, sc, st = C * V * T * H * W, T * H * W, V * T * H * W, H * W
        mean = (ctypes.c_float * 4)(*src.mean, 0.0)
        std = (ctypes.c_float * 4)(*src.std, 1.0)
    elif src.dim() == 6:
        t = src
        B, V, C, T, H, W = src.shape
        sb, sv, sc, st = V * C * T * H * W, C * T * H * W, T * H * W, H * W
    else:
        t = src
        B, C, T, H, W = src.shape
        V = 1
        sb, sv, sc, st = C * T * H * W, 0, T * H * W, H * W
    nv = V - first_view if n_views is None else n_views
    shape = (B * nv, T, H // 2, W // 2 + 3, 16) if s2d else (B * nv, T, H, W, 8)
    if fp32_mode():
        # split planes [K][clips][...]; ``out`` is a clip range of an ingest_buffer() (planes stay plane_stride apart)
        K = F32_PLANES
        if out is not None:
            assert tuple(out.shape) == (K,) + shape and out.dtype == torch.bfloat16 and out[0].is_contiguous()
            dst = out
        else:
            dst = torch.empty((K,) + shape, dtype=torch.bfloat16, device=t.device)
        call("dv_ingest_clips_planes", ptr(t), 1 if t.dtype == torch.uint8 else 0, ptr(dst), dst.stride(0), K, ptr(perm),
             sb, sv, sc, st, B, C, T, H, W, first_view, nv, n_series, mean, std, 1 if s2d else 0, stream_ptr())
        return Act(None, C, needs_grad=False, s2d=(B * nv, T, H, W) if s2d else None, planes=dst)
    if out is not None:
        assert tuple(out.shape) == shape and out.is_contiguous() and out.dtype == torch.bfloat16
        dst = out
    else:
        dst = torch.empty(shape, dtype=torch.bfloat16, device=t.device)
    call("dv_ingest_clips_u8" if t.dtype == torch.uint8 else "dv_ingest_clips", ptr(t), ptr(dst), ptr(perm), sb, sv, sc, st,
         B, C, T, H, W, first_view, nv,
         n_series, mean, std, 1 if s2d else 0, stream_ptr())
    return Act(dst, C, needs_grad=False, s2d=(B * nv, T, H, W) if s2d else None)


def ingest_shape(src, n_clips, s2d):
    """Shape of the ingest buffer for n_clips clips of ``src``'s frame geometry."""
    T, H, W = (src.block_shape if isinstance(src, RawClips) else tuple(src.shape))[-3:]
    return (n_clips, T, H // 2, W // 2 + 3, 16) if s2d else (n_clips, T, H, W, 8)


def ingest_buffer(src, n_clips, s2d, device):
    """Uninitialised ingest destination for n_clips clips (several ingest() calls fill clip ranges of it)."""
    shape = ingest_shape(src, n_clips, s2d)
    if fp32_mode():
        shape = (F32_PLANES,) + shape
    return torch.empty(shape, dtype=torch.bfloat16, device=device)


def clip_range(buf, lo, hi):
    """Clips [lo, hi) of an ingest_buffer() as the ``out`` argument of ingest()."""
    return buf[:, lo:hi] if fp32_mode() else buf[lo:hi]


def input_act(buf, C, s2d_dims=None):
    """The (gradient-free) input activation over a filled ingest buffer."""
    if fp32_mode():
        return Act(None, C, needs_grad=False, s2d=s2d_dims, planes=buf)
    return Act(buf, C, needs_grad=False, s2d=s2d_dims)


def input_tensor(act):
    """The tensor that holds an input activation's clips (clip dimension: see clip_dim())."""
    return act.planes if act.data is None else act.data


def clip_dim():
    return 1 if fp32_mode() else 0


def conv_stats(ctx, x, conv, bn):
    """y = conv(x) (raw, bf16) with fused per-channel sum/sumsq, then BN finalize -> scale/shift.
    Reference: nn.Conv3d + the statistics half of nn.BatchNorm3d (e.g. backbone/r21d.py:68)."""
    w = conv.weight
    Cout, Cin = w.shape[0], w.shape[1]
    stem = x.s2d is not None
    if stem:
        N, T, H, W = x.s2d
    else:
        N, T, H, W, Cin_p = x.shape5
    g = make_geom(N, T, H, W, Cin, Cout, tuple(w.shape[2:]), tuple(conv.stride), tuple(conv.padding))
    dev = x.device
    training_stats = ctx.training and bn.training if bn is not None else False
    stats = ctx.zeros64(2 * g.Cout_p, dev) if training_stats else None
    if stem:
        assert stem_eligible(conv, H, W), "space-to-depth input needs a stride-2 7x7 first conv"
    else:
        assert g.Cin_p == Cin_p, (g.Cin_p, Cin_p)
    if fp32_mode():
        # y (fp32) = sum of the plane products; the batch statistics are a separate pass over the finished sum
        y = torch.empty((N, g.To, g.Ho, g.Wo, g.Cout_p), dtype=torch.float32, device=dev)
        packed = packed_stem_weight_planes(conv, g) if stem else packed_weight_planes(conv)
        bias = _bias_padded(conv, g.Cout_p)
        # (a strided layer reads every plane through one tensor map per stride-parity class its taps reach: 12 maps at most)
        n_parity = 1
        for k_, s_ in zip(w.shape[2:], conv.stride):
            n_parity *= min(int(k_), int(s_))
        if F32_MERGE and not stem and n_parity * F32_PLANES <= 12:
            wf_all, _ = packed_weight_planes_all(conv)      # every plane product in one launch, y stored once
            call("dv_conv3d_fprop_f32planes", ptr(x.planes), x.planes.stride(0), F32_PLANES, ptr(wf_all), ptr(y),
                 ptr(stats), ptr(bias), ctypes.byref(g), stream_ptr())      # batch statistics from the epilogue
        else:
            for n, (i, j) in enumerate(_terms()):      # the first product overwrites y, the others add to it
                call("dv_conv3d_stem_fprop_f32acc" if stem else "dv_conv3d_fprop_f32acc", ptr(x.planes[i]),
                     ptr(packed[j][0]), ptr(y), ptr(bias) if n == 0 else None, ctypes.byref(g), 1 if n else 0,
                     stream_ptr())
            if training_stats:
                call("dv_f32_colstats", ptr(y), ptr(stats), y.numel() // g.Cout_p, g.Cout_p, stream_ptr())
    elif stem:
        y = torch.empty((N, g.To, g.Ho, g.Wo, g.Cout_p), dtype=torch.bfloat16, device=dev)
        packed = (packed_stem_weights(conv, g), None)
        call("dv_conv3d_stem_fprop_bf16", ptr(x.data), ptr(packed[0]), ptr(y), ptr(stats),
             ptr(_bias_padded(conv, g.Cout_p)), ctypes.byref(g), stream_ptr())
    elif x.lazy is not None:
        # consumer-side BatchNorm: this conv reads the RAW output of the conv below and applies its BatchNorm + ReLU
        # to the operand tile in shared memory (csrc/bn_xform.cuh) - the activation in between never exists in HBM
        y = torch.empty((N, g.To, g.Ho, g.Wo, g.Cout_p), dtype=torch.bfloat16, device=dev)
        packed = packed_weights(conv)
        rp, relu_p = x.lazy
        call("dv_conv3d_fprop_bnrelu_bf16", ptr(rp.y), ptr(rp.ss), 1 if relu_p else 0, ptr(packed[0]), ptr(y), ptr(stats),
             ptr(_bias_padded(conv, g.Cout_p)), ctypes.byref(g), stream_ptr())
    else:
        y = torch.empty((N, g.To, g.Ho, g.Wo, g.Cout_p), dtype=torch.bfloat16, device=dev)
        packed = packed_weights(conv)
        call("dv_conv3d_fprop_bf16", ptr(x.data), ptr(packed[0]), ptr(y), ptr(stats),
             ptr(_bias_padded(conv, g.Cout_p)), ctypes.byref(g), stream_ptr())
    r = RawBN()
    r.y, r.geom, r.x, r.conv, r.bn, r.packed, r.stem = y, g, x, conv, bn, packed, stem
    r.count = float(N * g.To * g.Ho * g.Wo)
    r.sync = False
    if bn is None:
        r.ss = r.saved = None
        return r
    r.ss = torch.empty(2 * g.Cout_p, dtype=torch.float32, device=dev)
    r.saved = torch.empty(2 * g.Cout_p, dtype=torch.float32, device=dev) if training_stats else None
    momentum = bn.momentum if bn.momentum is not None else 0.1
    track = bn.track_running_stats and bn.running_mean is not None
    # a pass that runs concurrently with another one (BackbonePairFunction) must not touch the running statistics
    # in place: its updates are applied after the other pass's, in the reference's order, at the join
    deferred = track and training_stats and ctx.defer_running is not None
    upd = track and not deferred
    fin_args = (ptr(stats), ptr(bn.weight.detach()), ptr(bn.bias.detach()),
                ptr(bn.running_mean) if upd else None, ptr(bn.running_var) if upd else None,
                ptr(r.ss), ptr(r.saved), Cout, g.Cout_p)
    if training_stats and _is_sync(bn):
        r.count *= dist.get_world_size()
        r.sync = True
        peer = comm.peer_state(dev, 2 * g.Cout_p)
        if peer is not None:      # exchange + finalisation in one launch over NVLink peer memory
            call("dv_bn_finalize_sync", *fin_args, ctypes.c_double(r.count), ctypes.c_float(bn.eps),
                 ctypes.c_float(momentum), *peer.next_call(), stream_ptr())
            _after_finalize(ctx, r, bn, momentum, upd, deferred)
            return r
        comm.small_allreduce_(stats)
    call("dv_bn_finalize", *fin_args, ctypes.c_double(r.count), ctypes.c_float(bn.eps),
         ctypes.c_float(momentum), 1 if training_stats else 0, stream_ptr())
    if training_stats:
        _after_finalize(ctx, r, bn, momentum, upd, deferred)
    return r


def _after_finalize(ctx, r, bn, momentum, upd, deferred):
    if upd:
        # one multi-tensor increment per pass (flush_batch_counters) instead of a torch kernel per BatchNorm layer
        ctx.nbt_pending.append(bn.num_batches_tracked)
    elif deferred:
        ctx.defer_running.append((bn, r.saved, r.count, momentum))


def flush_batch_counters(ctx):
    """num_batches_tracked += 1 for every BatchNorm layer the pass updated (nn.BatchNorm3d.forward), as one
    multi-tensor launch. A layer met twice in one pass is incremented twice."""
    if ctx.nbt_pending:
        uniq, counts = [], {}
        for t in ctx.nbt_pending:
            if id(t) not in counts:
                uniq.append(t)
            counts[id(t)] = counts.get(id(t), 0) + 1
        torch._foreach_add_(uniq, 1)
        for t in uniq:
            if counts[id(t)] > 1:
                t += counts[id(t)] - 1
        ctx.nbt_pending = []


def apply_deferred_running(items):
    """Running-statistic updates of a concurrent pass, in layer order on the current stream (nn.BatchNorm3d momentum
    update with the unbiased batch variance; mean / invstd come from the pass's saved statistics)."""
    for bn, saved, count, momentum in items:
        C = bn.running_mean.shape[0]
        Cp = saved.shape[0] // 2
        mean, invstd = saved[:C], saved[Cp:Cp + C]
        var = (1.0 / (invstd.double() * invstd.double()) - bn.eps).clamp_min(0.0)
        unbiased = (var * (count / (count - 1.0) if count > 1.0 else 1.0)).float()
        bn.running_mean.mul_(1.0 - momentum).add_(mean, alpha=momentum)
        bn.running_var.mul_(1.0 - momentum).add_(unbiased, alpha=momentum)
        bn.num_batches_tracked += 1


# The weight gradient of a layer is off the backward critical path (dgrad -> BN backward of the layer below -> ...),
# so it runs on a side stream: its tensor-bound CTAs then share the SMs with the HBM-bound BatchNorm passes of the
# layers below instead of running back to back with them. DV_WGRAD_STREAM=0 keeps everything on one stream.
WGRAD_SIDE_STREAM = os.environ.get("DV_WGRAD_STREAM", "1") != "0"
_side_streams = {}


def _side_stream(device):
    """The weight-gradient stream that belongs to the current (issuing) stream."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    s = _side_streams.get(key)
    if s is None:
        s = _side_streams[key] = torch.cuda.Stream(device=device)
    return s


_DIAG_SKIP_WGRAD = os.environ.get("DV_DIAG_SKIP_WGRAD", "0") != "0"   # timing experiments only: gradients are WRONG


def _wgrad(r, dy, gw, beta=0.0):
    g = r.geom
    if _DIAG_SKIP_WGRAD:
        gw.zero_()
        return
    if r.stem:
        dwp = torch.empty((g.Cout_p, g.kt * 4, 64), dtype=torch.float32, device=dy.device)
        call("dv_conv3d_stem_wgrad_bf16", ptr(r.x.data), ptr(dy), ptr(dwp), ctypes.byref(g), stream_ptr())
        call("dv_unpack_stem_wgrad", ptr(dwp), ptr(gw), ctypes.byref(g), ctypes.c_float(beta), stream_ptr())
    else:
        dwp = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.float32, device=dy.device)
        if r.x.lazy is not None:      # the operand z is recomputed from the raw y of the conv below (never stored)
            rp, relu_p = r.x.lazy
            call("dv_conv3d_wgrad_bnrelu_bf16", ptr(rp.y), ptr(rp.ss), 1 if relu_p else 0, ptr(dy), ptr(dwp),
                 ctypes.byref(g), stream_ptr())
        else:
            call("dv_conv3d_wgrad_bf16", ptr(r.x.data), ptr(dy), ptr(dwp), ctypes.byref(g), stream_ptr())
        call("dv_unpack_conv_wgrad", ptr(dwp), ptr(gw), ctypes.byref(g), ctypes.c_float(beta), stream_ptr())


def _conv_backward_f32(ctx, r, dyp):
    """fp32 mode: dyp = split planes [K, ...] of dy. Weight and data gradient as sums of plane products."""
    g = r.geom
    gw = torch.empty_like(r.conv.weight)
    terms = _terms()
    if r.stem:
        dwp = torch.empty((g.Cout_p, g.kt * 4, 64), dtype=torch.float32, device=dyp.device)
        for n, (i, j) in enumerate(terms):
            call("dv_conv3d_stem_wgrad_bf16" if n == 0 else "dv_conv3d_stem_wgrad_bf16_acc", ptr(r.x.planes[i]),
                 ptr(dyp[j]), ptr(dwp), ctypes.byref(g), stream_ptr())
        call("dv_unpack_stem_wgrad", ptr(dwp), ptr(gw), ctypes.byref(g), ctypes.c_float(0.0), stream_ptr())
    else:
        dwp = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.float32, device=dyp.device)
        if F32_MERGE and r.x.planes.is_contiguous() and dyp.is_contiguous():
            call("dv_conv3d_wgrad_f32planes", ptr(r.x.planes), ptr(dyp), F32_PLANES, ptr(dwp), ctypes.byref(g), stream_ptr())
        else:
            for n, (i, j) in enumerate(terms):
                call("dv_conv3d_wgrad_bf16" if n == 0 else "dv_conv3d_wgrad_bf16_acc", ptr(r.x.planes[i]), ptr(dyp[j]),
                     ptr(dwp), ctypes.byref(g), stream_ptr())
        call("dv_unpack_conv_wgrad", ptr(dwp), ptr(gw), ctypes.byref(g), ctypes.c_float(0.0), stream_ptr())
    ctx.add_param_grad(r.conv.weight, gw)
    if r.conv.bias is not None:
        ctx.add_param_grad(r.conv.bias, torch.zeros_like(r.conv.bias))
    if r.x.needs_grad:
        dx = torch.empty(r.x.shape5, dtype=torch.float32, device=dyp.device)
        if F32_MERGE:
            _, wt_all = packed_weight_planes_all(r.conv)
            call("dv_conv3d_dgrad_f32planes", ptr(dyp), dyp.stride(0), F32_PLANES, ptr(wt_all), ptr(dx), ctypes.byref(g),
                 stream_ptr())
        else:
            for n, (i, j) in enumerate(terms):
                call("dv_conv3d_dgrad_f32acc", ptr(dyp[i]), ptr(r.packed[j][1]), ptr(dx), ctypes.byref(g), 1 if n else 0,
                     stream_ptr())
        _acc_grad(r.x, dx)


def _conv_backward(ctx, r, dy):
    """wgrad into the parameter-gradient sink, dgrad into r.x.grad."""
    g = r.geom
    first = id(r.conv.weight) not in ctx.param_grads
    direct = ctx.direct_target(r.conv.weight)
    beta = 0.0
    if direct is not None:      # added straight into .grad (zeroed by zero_grad; the other pass of the step adds too)
        gw, beta, first = direct, 1.0, False
        ctx.direct.add(id(r.conv.weight))
    else:
        gw = ctx.grad_out(r.conv.weight) if first else torch.empty_like(r.conv.weight)
    if WGRAD_SIDE_STREAM:
        main = torch.cuda.current_stream()
        side = _side_stream(dy.device)
        side.wait_stream(main)                 # dy (and, on the first use, x) are ready
        with torch.cuda.stream(side):
            _wgrad(r, dy, gw, beta)
        # the caching allocator must not hand these blocks to later main-stream work while the side stream reads them
        dy.record_stream(side)
        r.x.stored.record_stream(side)
        gw.record_stream(side)
        ctx.side_used = True
    else:
        side = None
        _wgrad(r, dy, gw, beta)
    if direct is None:
        ctx.add_param_grad(r.conv.weight, gw)
    if first:
        ctx.grad_ready(r.conv.weight, side)
    if r.conv.bias is not None:
        # a bias in front of training-mode BN has exactly zero gradient (BN removes the mean)
        ctx.add_param_grad(r.conv.bias, torch.zeros_like(r.conv.bias))
    if r.x.needs_grad:
        dx = torch.empty_like(r.x.stored)
        if r.x.bnred is not None and r.x.grad is None and _fuse_reduce_pays(g):
            # x = relu?(BN(y_prev)) and this is (so far) its only gradient: the dgrad epilogue also produces
            # the BN-backward sums of the layer below, saving dv_bn_bwd_reduce's pass over dx and y_prev
            rp, relu = r.x.bnred
            sums = ctx.zeros64(2 * g.Cin_p, dy.device)
            call("dv_conv3d_dgrad_bnred_bf16", ptr(dy), ptr(r.packed[1]), ptr(dx), ctypes.byref(g), ptr(rp.y),
                 ptr(rp.ss) if relu else None, ptr(sums), stream_ptr())
            r.x.fused = (sums, dx)
        else:
            call("dv_conv3d_dgrad_bf16", ptr(dy), ptr(r.packed[1]), ptr(dx), ctypes.byref(g), stream_ptr())
        _acc_grad(r.x, dx)


def _bn_bwd_finalize(ctx, r, sums, Cp, dev):
    """Backward sums of one BatchNorm -> dgamma / dbeta (into the gradient sink) and the coefficients of
    dy = A*g + B*y + C; cross-replica BatchNorm exchanges the sums first (fused over NVLink peer memory)."""
    bn = r.bn
    first = id(bn.weight) not in ctx.param_grads
    dw_direct, db_direct = ctx.direct_target(bn.weight), ctx.direct_target(bn.bias)
    direct = dw_direct is not None and db_direct is not None
    grad_beta = 0.0
    if direct:                  # dgamma / dbeta are added straight into .grad (see direct_param_grads)
        dgamma, dbeta, grad_beta, first = dw_direct, db_direct, 1.0, False
        ctx.direct.add(id(bn.weight))
        ctx.direct.add(id(bn.bias))
    else:
        dgamma = ctx.grad_out(bn.weight) if first else torch.empty_like(bn.weight)
        dbeta = ctx.grad_out(bn.bias) if first else torch.empty_like(bn.bias)
    coef = torch.empty(3 * Cp, dtype=torch.float32, device=dev)
    peer = comm.peer_state(dev, 2 * Cp) if r.sync else None
    if peer is not None:  # exchange of the sums + finalisation in one launch over NVLink peer memory
        call("dv_bn_bwd_finalize_sync", ptr(sums), ptr(bn.weight.detach()), ptr(r.saved), ptr(dgamma),
             ptr(dbeta), ptr(coef), r.geom.Cout, Cp, ctypes.c_double(r.count), ctypes.c_float(grad_beta),
             *peer.next_call(), stream_ptr())
    else:
        sums_g = sums
        if r.sync:
            sums_g = sums.clone()
            comm.small_allreduce_(sums_g)
        call("dv_bn_bwd_finalize", ptr(sums), ptr(sums_g), ptr(bn.weight.detach()), ptr(r.saved),
             ptr(dgamma), ptr(dbeta), ptr(coef), r.geom.Cout, Cp, ctypes.c_double(r.count),
             ctypes.c_float(grad_beta), stream_ptr())
    if not direct:
        ctx.add_param_grad(bn.weight, dgamma)
        ctx.add_param_grad(bn.bias, dbeta)
    if first:
        ctx.grad_ready(bn.weight)
        ctx.grad_ready(bn.bias)
    return coef


def _activate_f32(ctx, r1, r2, res, relu, out=None, out_coff=0):
    """fp32 mode of activate(): out = relu?(BN(r1) [+ BN(r2)] [+ res]) as fp32 plus its split planes, dense or into
    the channel slice [out_coff, out_coff + Cp) of the concat activation ``out``."""
    g = r1.geom
    Cp = g.Cout_p
    dev = r1.y.device
    K = F32_PLANES
    rows = r1.y.numel() // Cp
    if out is None:
        out_t = torch.empty_like(r1.y)
        planes = torch.empty((K,) + tuple(r1.y.shape), dtype=torch.bfloat16, device=dev)
        out_act = Act(out_t, g.Cout, planes=planes)
        out_ld = Cp
    else:
        out_act, out_t, planes, out_ld = out, out.data, out.planes, out.Cp
    call("dv_f32_bn_apply", ptr(r1.y), ptr(r1.ss), ptr(r2.y) if r2 else None, ptr(r2.ss) if r2 else None,
         ptr(res.data) if res is not None else None, ptr(out_t), ptr(planes), planes.stride(0), K, rows, Cp, out_ld,
         out_coff, 1 if relu else 0, stream_ptr())
    if not ctx.record:
        return out_act

    def backward():
        dout2 = None
        if out is None:
            dout, dout2 = out_act.grad, out_act.grad2
        else:
            dout = _materialize_grad(out_act)
        o_ld, o_coff = out_ld, out_coff
        ov = ctx.overrides.pop((id(out_act), out_coff), None)
        if ov is not None:      # a gate in front of this slice already produced the dense gradient
            dout, dout2, o_ld, o_coff = ov, None, Cp, 0
        assert dout is not None, "activation has no gradient"
        need_g = res is not None and res.needs_grad
        g_buf = None
        mask_ss = ptr(r1.ss) if (relu and r2 is None and res is None) else None
        assert ov is None or mask_ss is not None
        for r in (r1, r2):
            if r is None:
                continue
            sums = ctx.zeros64(2 * Cp, dev)
            call("dv_f32_bn_bwd_reduce", ptr(dout), ptr(dout2), ptr(out_t), ptr(r.y), mask_ss, ptr(sums), rows, Cp,
                 o_ld, o_coff, 1 if relu else 0, stream_ptr())
            coef = _bn_bwd_finalize(ctx, r, sums, Cp, dev)
            dyp = torch.empty((K,) + tuple(r.y.shape), dtype=torch.bfloat16, device=dev)
            want_g = need_g and g_buf is None
            if want_g:
                g_buf = torch.empty_like(r.y)
            call("dv_f32_bn_bwd_apply", ptr(dout), ptr(dout2), ptr(out_t), ptr(r.y), mask_ss, ptr(coef), ptr(dyp),
                 dyp.stride(0), K, ptr(g_buf) if want_g else None, rows, Cp, o_ld, o_coff, 1 if relu else 0,
                 stream_ptr())
            _conv_backward_f32(ctx, r, dyp)
        if need_g:
            _acc_grad(res, g_buf)
        if out is None:
            out_act.grad = out_act.grad2 = None

    ctx.tape.append(backward)
    return out_act


# Consumer-side BatchNorm (csrc/bn_xform.cuh): a plain relu?(BN(y)) whose only consumer is a convolution is not
# written to HBM - the consumer's fprop (and wgrad) load the raw y and normalise the operand tile in shared memory.
# Results are bit-identical to the stand-alone dv_bn_apply pass. Measured on B200 (profiles/r02_bn_fusion_*.txt): the
# fused fprop costs +0.3-0.4 ms where the pass it replaces costs 0.6-1.0 ms - a win - but the fused wgrad costs +0.6-1.1 ms
# with nothing more to save (the transform's shared-memory traffic competes with the tensor core's operand fetch), so a
# TRAINING step is slower with it. Hence: 1 (default) = forward-only passes (evaluation, feature extraction, MoCo's key
# encoder), 2 = training passes too, 0 = never.
FUSE_BN_APPLY = int(os.environ.get("DV_FUSE_BN_APPLY", "1"))


def activate(ctx, r1, r2=None, res=None, relu=True, out=None, out_coff=0, conv_only=False):
    """out = relu?(BN(r1) [+ BN(r2)] [+ res]). ``out``/``out_coff`` let several branches write
    channel slices of one tensor (concat-free Inception, backbone/s3dg.py:130). ``conv_only``: the caller promises
    that the result feeds exactly one conv_stats() and nothing else (the BatchNorm + ReLU between the two convs of a
    factorised convolution) - it is then applied inside that convolution instead of being stored.
    Reference: BatchNorm3d affine + ReLU + residual add (backbone/r21d.py:56-57,116-122)."""
    if fp32_mode():
        return _activate_f32(ctx, r1, r2, res, relu, out, out_coff)
    g = r1.geom
    Cp = g.Cout_p
    dev = r1.y.device
    lazy = conv_only and r2 is None and res is None and out is None and \
        (FUSE_BN_APPLY == 2 or (FUSE_BN_APPLY == 1 and not ctx.record))
    if lazy:
        out_t = None
        out_act = Act(None, g.Cout, lazy=(r1, relu))
        out_ld = Cp
    elif out is None:
        out_t = torch.empty_like(r1.y)
        out_act = Act(out_t, g.Cout)
        out_ld = Cp
    else:
        out_act = out
        out_t = out.data
        out_ld = out.Cp
    rows = r1.y.numel() // Cp
    if not lazy:
        call("dv_bn_apply", ptr(r1.y), ptr(r1.ss), ptr(r2.y) if r2 else None, ptr(r2.ss) if r2 else None,
             ptr(res.data) if res is not None else None, ptr(out_t), rows, Cp, out_ld, out_coff,
             1 if relu else 0, stream_ptr())
    if not ctx.record:
        return out_act
    plain = r2 is None and res is None and out is None
    if plain:
        out_act.bnred = (r1, relu)

    def backward():
        dout2 = None
        if out is None:
            dout, dout2 = out_act.grad, out_act.grad2      # two pending contributions are summed in-kernel
        else:
            dout = _materialize_grad(out_act)
        o_ld, o_coff = out_ld, out_coff
        ov = ctx.overrides.pop((id(out_act), out_coff), None)
        if ov is not None:      # a gate in front of this slice already produced the dense gradient
            dout, dout2, o_ld, o_coff = ov, None, Cp, 0
        assert dout is not None, "activation has no gradient"
        need_g = res is not None and res.needs_grad
        g_buf = None
        # plain relu(BN(y)): the mask is recomputed from y, `out` is not re-read
        mask_ss = ptr(r1.ss) if (relu and r2 is None and res is None) else None
        assert ov is None or mask_ss is not None
        for i, r in enumerate((r1, r2)):
            if r is None:
                continue
            fused = out_act.fused if plain else None
            if fused is not None and dout2 is None and ov is None and fused[1] is dout:
                sums = fused[0]      # already reduced by the dgrad that produced dout
            else:
                sums = ctx.zeros64(2 * Cp, dev)
                call("dv_bn_bwd_reduce", ptr(dout), ptr(dout2), ptr(out_t), ptr(r.y), mask_ss, ptr(sums), rows, Cp,
                     o_ld, o_coff, 1 if relu else 0, stream_ptr())
            coef = _bn_bwd_finalize(ctx, r, sums, Cp, dev)
            dy = torch.empty_like(r.y)
            want_g = need_g and g_buf is None
            if want_g:
                g_buf = torch.empty_like(r.y)
            call("dv_bn_bwd_apply", ptr(dout), ptr(dout2), ptr(out_t), ptr(r.y), mask_ss, ptr(coef), ptr(dy),
                 ptr(g_buf) if want_g else None, rows, Cp, o_ld, o_coff, 1 if relu else 0, stream_ptr())
            _conv_backward(ctx, r, dy)
        if need_g:
            _acc_grad(res, g_buf)
        if out is None:
            out_act.grad = out_act.grad2 = None
        out_act.fused = None

    ctx.tape.append(backward)
    return out_act


def new_concat(x, C_total):
    """Empty activation with x's N,T,H,W that several branches fill slice by slice
    (torch.cat of the Inception branches, backbone/s3dg.py:130)."""
    N, T, H, W, _ = x.shape5
    assert C_total % 8 == 0
    if fp32_mode():
        dev = x.device
        return Act(torch.empty((N, T, H, W, C_total), dtype=torch.float32, device=dev), C_total,
                   planes=torch.empty((F32_PLANES, N, T, H, W, C_total), dtype=torch.bfloat16, device=dev))
    return Act(torch.empty((N, T, H, W, C_total), dtype=torch.bfloat16, device=x.data.device), C_total)


def self_gate(ctx, cat, coff, raw, fc):
    """S3D-G SelfGating on the slice [coff, coff+C) of ``cat`` that ``raw``'s activation just wrote:
    w = sigmoid(fc(mean_{T,H,W} z)); slice *= w (backbone/s3dg.py:68-78)."""
    N, T, H, W, ld = cat.shape5
    S = T * H * W
    C, Cp = raw.geom.Cout, raw.geom.Cout_p
    dev = cat.data.device
    f32 = cat.data.dtype == torch.float32      # fp32 mode: same flow on the fp32 kernels, planes kept in step
    pre_ = "dv_f32_" if f32 else "dv_"
    mean = torch.empty((N, C), dtype=torch.float32, device=dev)
    call(pre_ + "slice_mean", ptr(cat.data), ptr(mean), N, S, C, ld, coff, stream_ptr())
    w = torch.empty((N, C), dtype=torch.float32, device=dev)
    call("dv_gate_fc_fwd", ptr(mean), ptr(fc.weight.detach()), ptr(fc.bias.detach()), ptr(w), N, C, stream_ptr())
    if f32:
        call("dv_f32_gate_scale", ptr(cat.data), ptr(cat.planes), cat.planes.stride(0), F32_PLANES, ptr(w), N, S, C, ld,
             coff, stream_ptr())
    else:
        call("dv_gate_scale", ptr(cat.data), ptr(w), N, S, C, ld, coff, stream_ptr())
    if not ctx.record:
        return

    def backward():
        dout = _materialize_grad(cat)
        dw = torch.empty((N, C), dtype=torch.float32, device=dev)
        call(pre_ + "gate_bwd_reduce", ptr(dout), ptr(raw.y), ptr(raw.ss), ptr(dw), N, S, C, Cp, ld, coff, stream_ptr())
        dpre = torch.empty_like(dw)
        gW = torch.empty_like(fc.weight)
        gb = torch.empty_like(fc.bias)
        dmean = torch.empty_like(dw)
        call("dv_gate_fc_bwd", ptr(dw), ptr(w), ptr(mean), ptr(fc.weight.detach()), ptr(dpre), ptr(gW), ptr(gb), ptr(dmean),
             N, C, stream_ptr())
        ctx.add_param_grad(fc.weight, gW)
        ctx.add_param_grad(fc.bias, gb)
        dz = torch.empty((N, T, H, W, Cp), dtype=cat.data.dtype, device=dev)
        call(pre_ + "gate_bwd_apply", ptr(dout), ptr(w), ptr(dmean), ptr(dz), N, S, C, Cp, ld, coff, stream_ptr())
        ctx.overrides[(id(cat), coff)] = dz

    ctx.tape.append(backward)


def max_pool(ctx, x, kernel, stride, padding):
    """nn.MaxPool3d (backbone/c3d.py:18, backbone/s3dg.py:151)."""
    N, T, H, W, Cp = x.shape5
    kt, kh, kw = kernel
    st, sh, sw = stride
    pt, ph, pw = padding
    To, Ho, Wo = (T + 2 * pt - kt) // st + 1, (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    geom = (ctypes.c_int32 * 17)(N, T, H, W, To, Ho, Wo, Cp, kt, kh, kw, st, sh, sw, pt, ph, pw)
    if fp32_mode():
        return _max_pool_f32(ctx, x, geom, (N, To, Ho, Wo, Cp))
    y = torch.empty((N, To, Ho, Wo, Cp), dtype=torch.bfloat16, device=x.data.device)
    track = ctx.record and x.needs_grad
    # training: 1-byte argmax per output element (backward becomes a gather instead of re-scanning windows for ties,
    # which is what made the 3x3x3 stride-1 pools of the Inception blocks 20 % of an S3D-G step)
    idx = torch.empty((N, To, Ho, Wo, Cp), dtype=torch.uint8, device=x.data.device) if track else None
    if track:
        call("dv_maxpool3d_fwd_idx", ptr(x.data), ptr(y), ptr(idx), geom, stream_ptr())
    else:
        call("dv_maxpool3d_fwd", ptr(x.data), ptr(y), geom, stream_ptr())
    out = Act(y, x.C)
    if track:
        def backward():
            dx = torch.empty_like(x.data)
            call("dv_maxpool3d_bwd_idx", ptr(idx), ptr(_materialize_grad(out)), ptr(dx), geom, stream_ptr())
            _acc_grad(x, dx)
            out.grad = None
        ctx.tape.append(backward)
    return out


def _max_pool_f32(ctx, x, geom, oshape):
    dev = x.data.device
    K = F32_PLANES
    y = torch.empty(oshape, dtype=torch.float32, device=dev)
    planes = torch.empty((K,) + tuple(oshape), dtype=torch.bfloat16, device=dev)
    track = ctx.record and x.needs_grad
    idx = torch.empty(oshape, dtype=torch.uint8, device=dev) if track else None
    call("dv_f32_maxpool3d_fwd", ptr(x.data), ptr(y), ptr(idx), ptr(planes), planes.stride(0), K, geom, stream_ptr())
    out = Act(y, x.C, planes=planes)
    if track:
        def backward():
            dx = torch.empty_like(x.data)
            call("dv_f32_maxpool3d_bwd", ptr(idx), ptr(_materialize_grad(out)), ptr(dx), geom, stream_ptr())
            _acc_grad(x, dx)
            out.grad = None
        ctx.tape.append(backward)
    return out


def global_pool(ctx, x):
    """AdaptiveAvgPool3d((1,1,1)) -> fp32 [N, C] (model/simclr.py:166)."""
    N, T, H, W, Cp = x.shape5
    S = T * H * W
    out = torch.empty((N, x.C), dtype=torch.float32, device=x.data.device)
    call("dv_f32_avgpool_fwd" if x.data.dtype == torch.float32 else "dv_avgpool_fwd", ptr(x.data), ptr(out), N, S, x.C, Cp, x.C, stream_ptr())
    return out


def global_pool_backward(x, dpooled):
    N, T, H, W, Cp = x.shape5
    dx = torch.empty_like(x.data)
    d = dpooled.contiguous()
    call("dv_f32_avgpool_bwd" if dx.dtype == torch.float32 else "dv_avgpool_bwd", ptr(d), ptr(dx), N, T * H * W, x.C, Cp, x.C, stream_ptr())
    _acc_grad(x, dx)


def to_ncdhw(x):
    N, T, H, W, Cp = x.shape5
    out = torch.empty((N, x.C, T, H, W), dtype=torch.float32, device=x.data.device)
    call("dv_f32_ndhwc_to_ncdhw" if x.data.dtype == torch.float32 else "dv_ndhwc_bf16_to_ncdhw", ptr(x.data), ptr(out),
         N, x.C, Cp, T * H * W, stream_ptr())
    return out


def from_ncdhw_grad(x, d):
    N, T, H, W, Cp = x.shape5
    dx = torch.empty_like(x.data)
    d = d.contiguous()
    call("dv_f32_ncdhw_to_ndhwc" if dx.dtype == torch.float32 else "dv_ncdhw_to_ndhwc_bf16", ptr(d), ptr(dx), N, x.C, Cp,
         T * H * W, stream_ptr())
    _acc_grad(x, dx)


def run_backward(ctx):
    for fn in reversed(ctx.tape):
        fn()
    ctx.tape.clear()
    if ctx.side_used:      # join: parameter gradients are consumed on the current stream from here on
        dev = torch.cuda.current_device()
        torch.cuda.current_stream().wait_stream(_side_stream(torch.device("cuda", dev)))
        ctx.side_used = False


# ----------------------------------------------------------------------------- autograd bridge
class BackboneFunction(torch.autograd.Function):
    """One autograd node for a whole backbone pass. ``program(ctx, x_act) -> Act`` builds the graph with
    the primitives above; parameters are passed as inputs so that their gradients flow through the
    normal autograd accumulation (DDP hooks, optimizers and .grad all keep working)."""

    @staticmethod
    def forward(fctx, program, make_input, pooled, training, record, reducer, *params):
        ectx = Context(training, record=record)
        x = make_input()
        feat = program(ectx, x)
        flush_batch_counters(ectx)
        fctx.ectx, fctx.feat, fctx.pooled, fctx.params, fctx.reducer = ectx, feat, pooled, params, reducer
        fctx.set_materialize_grads(False)
        out = global_pool(ectx, feat) if pooled else to_ncdhw(feat)
        if not record:
            fctx.feat = None
        return out

    @staticmethod
    def backward(fctx, dout):
        ectx, feat = fctx.ectx, fctx.feat
        if dout is None or feat is None:
            return (None,) * (6 + len(fctx.params))
        # data-parallel training (parallel.DataParallel): parameter gradients are written into flat buckets and every
        # bucket is all-reduced on a communication stream as soon as it is complete, under the rest of this backward
        rp = ectx.rpass = fctx.reducer.begin_pass() if fctx.reducer is not None else None
        if fctx.pooled:
            global_pool_backward(feat, dout)
        else:
            from_ncdhw_grad(feat, dout)
        run_backward(ectx)
        if rp is not None:
            for p in fctx.params:       # gradients produced outside the bucket views (gating fc, zero biases, fp32 mode)
                g = ectx.param_grads.get(id(p))
                if g is not None and rp.manages(p) and id(p) not in rp.seen:
                    v = rp.view(p)
                    v.copy_(g)
                    ectx.param_grads[id(p)] = v
                    rp.ready(p)
            rp.finish()
            ectx.rpass = None
        grads = tuple(ectx.param_grads.get(id(p)) if p.requires_grad else None for p in fctx.params)
        ectx.param_grads = {}
        ectx.direct = set()
        fctx.feat = None
        return (None, None, None, None, None, None) + grads


# Two independent passes of one backbone (SimCLR+DualVar: the 3B clips and the B segment-shuffled clips,
# model/simclr.py:352-387) issued on two streams from ONE autograd node: the tensor-bound conv kernels of one pass
# then share the GPU with the HBM-bound BatchNorm passes of the other, forward and backward. The second pass lives
# entirely on its own stream (its activations are allocated, used and freed there); what crosses streams is ordered
# explicitly: input frames (event), packed weights (pack events), running statistics (deferred to the join, so the
# reference's update order pass 1 -> pass 2 is kept), outputs and parameter gradients (joins + record_stream).
# Measured on B200 (tests/diag/overlap_probe.py, profiles/r01b_stream_overlap.txt): a conv_tile_kernel launch and a
# BatchNorm pass on two streams take LONGER together than back to back (fprop 0.71 + bn_apply 0.49 ms: 1.18 sequential,
# 1.33-1.37 concurrent) - the persistent conv CTAs are themselves bound by the memory system (L2 -> SM feed), so the
# BatchNorm blocks that become resident next to them only slow both down. The whole step: 80.5 ms with the two passes
# on two streams vs 77.8 ms sequential. Results are identical to the sequential issue (tests/diag/pass_streams_check.py),
# so the path stays available behind DV_PASS_STREAMS=1, but it is off by default.
PASS_STREAMS = os.environ.get("DV_PASS_STREAMS", "0") != "0"
_pass_streams = {}


def _pass_stream(device):
    s = _pass_streams.get(device)
    if s is None:
        s = _pass_streams[device] = torch.cuda.Stream(device=device)
    return s


class BackbonePairFunction(torch.autograd.Function):
    @staticmethod
    def forward(fctx, program, make_inputs, training, record, *params):
        main = torch.cuda.current_stream()
        dev = params[0].device
        sb = _pass_stream(dev)
        start = main.record_event()
        ectx_a = Context(training, record=record)
        xa = make_inputs[0]()
        feat_a = program(ectx_a, xa)
        flush_batch_counters(ectx_a)
        out_a = global_pool(ectx_a, feat_a)
        ectx_b = Context(training, record=record)
        ectx_b.defer_running = []
        sb.wait_event(start)
        with torch.cuda.stream(sb):
            xb = make_inputs[1]()
            feat_b = program(ectx_b, xb)
            out_b = global_pool(ectx_b, feat_b)
        main.wait_stream(sb)
        out_b.record_stream(main)
        apply_deferred_running(ectx_b.defer_running)
        ectx_b.defer_running = None
        fctx.state = (ectx_a, feat_a if record else None, ectx_b, feat_b if record else None, sb)
        fctx.params = params
        fctx.set_materialize_grads(False)
        return out_a, out_b

    @staticmethod
    def backward(fctx, da, db):
        ectx_a, feat_a, ectx_b, feat_b, sb = fctx.state
        n_lead = 4
        if feat_a is None:
            return (None,) * (n_lead + len(fctx.params))
        main = torch.cuda.current_stream()
        ready = main.record_event()
        if da is not None:
            global_pool_backward(feat_a, da)
            run_backward(ectx_a)
        if db is not None:
            sb.wait_event(ready)
            db.record_stream(sb)
            with torch.cuda.stream(sb):
                global_pool_backward(feat_b, db)
                run_backward(ectx_b)
            main.wait_stream(sb)
        grads = []
        for p in fctx.params:
            ga = ectx_a.param_grads.get(id(p)) if p.requires_grad else None
            gb = ectx_b.param_grads.get(id(p)) if p.requires_grad else None
            if gb is not None:
                gb.record_stream(main)
            grads.append(gb if ga is None else (ga if gb is None else ga.add_(gb)))
        ectx_a.param_grads, ectx_b.param_grads = {}, {}
        fctx.state = None
        return (None,) * n_lead + tuple(grads)


def run_backbone_pair(module, program, make_inputs):
    """Pooled outputs of two independent passes of ``module`` (see BackbonePairFunction)."""
    params = [p for p in module.parameters()]
    record = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return BackbonePairFunction.apply(program, make_inputs, module.training, record, *params)


def run_backbone(module, program, make_input, pooled):
    params = [p for p in module.parameters()]
    record = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return BackboneFunction.apply(program, make_input, pooled, module.training, record,
                                  getattr(module, "_dv_reducer", None), *params)
