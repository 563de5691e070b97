"""Data-parallel wrapper with bucketed, overlapped gradient all-reduce (BASELINE north_star: "gradients are allreduced
with bucket overlap"; reference: torch.nn.parallel.DistributedDataParallel at pretrain.py:248).

Why not torch's DDP alone: a backbone pass of this package is ONE autograd node (engine.BackboneFunction), so DDP's
reducer sees every backbone gradient only when the whole backward has finished and its ~57 MB all-reduce cannot overlap
anything. ``DataParallel`` below keeps DDP's contract (same constructor position in the training script, ``.module``,
gradients averaged over ranks before ``optimizer.step()``, parameters and buffers broadcast from rank 0 at construction)
and moves the backbone gradients' reduction INTO the engine's backward:

* every encoder module gets a ``BucketedReducer``; the engine's backward (engine.run_backward) writes each weight /
  BatchNorm gradient straight into a flat per-pass buffer laid out in reverse parameter order and reports it ready;
* a bucket (~8 MB of consecutive parameters) is all-reduced (NCCL, averaged) on a communication stream as soon as its
  last gradient exists, while the layers below are still computing their dgrad / wgrad;
* a parameter used by several passes of one step (SimCLR+DualVar runs the encoder twice, MoCo three times) is reduced
  once per pass - all-reduce is linear, autograd then sums the reduced contributions exactly as it sums local ones;
* gradients that do not come out of the engine (projection heads, classifier tail: ~1 M parameters) are reduced in one
  flat call from an end-of-backward callback.

``no_sync()`` skips all reductions (gradient accumulation / attribution runs). Works with ``nn.SyncBatchNorm``-converted
models exactly like DDP does. On the gloo backend (CPU tests of the host logic) the reduction is SUM followed by a
division; on NCCL it is ReduceOp.AVG.
"""
import contextlib

import torch
import torch.distributed as dist
import torch.nn as nn


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _avg_(flat, group=None):
    """In-place average over ranks."""
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(dist.get_world_size(group))


class _Pass:
    """Flat gradient buffer of one backward pass of one encoder and the state of its buckets."""

    def __init__(self, owner):
        self.o = owner
        # unused parameters contribute zeros (and the alignment padding stays defined)
        self.flat = torch.zeros(owner.total, dtype=torch.float32, device=owner.device)
        self.pending = list(owner.bucket_sizes)       # gradients still missing per bucket
        self.launched = [False] * len(owner.buckets)
        self.events = {}                              # bucket -> {producer stream: event after its last gradient}
        self.seen = set()
        self.extra = []                               # (param, later contribution) once p's bucket is in flight

    def manages(self, p):
        return id(p) in self.o.slot

    def view(self, p):
        """The slice of the flat buffer that receives p's gradient (shaped like p), or None if p is not managed."""
        slot = self.o.slot.get(id(p))
        if slot is None:
            return None
        off, n, _ = slot
        return self.flat[off:off + n].view(p.shape)

    def _note(self, b, stream):
        if self.o.comm_stream is None:
            return
        st = stream if stream is not None else torch.cuda.current_stream(self.o.device)
        self.events.setdefault(b, {})[st.cuda_stream] = st.record_event()

    def ready(self, p, stream=None):
        """p's gradient has been written into view(p) by work queued on ``stream`` (default: the current stream)."""
        slot = self.o.slot.get(id(p))
        if slot is None or id(p) in self.seen:
            return
        self.seen.add(id(p))
        b = slot[2]
        self._note(b, stream)
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def in_flight(self, p):
        slot = self.o.slot.get(id(p))
        return slot is not None and self.launched[slot[2]]

    def _launch(self, b):
        o = self.o
        lo, hi = o.buckets[b]
        self.launched[b] = True
        if o.comm_stream is None:                      # CPU / gloo (host-logic tests): synchronous
            _avg_(self.flat[lo:hi], o.group)
            return
        for ev in self.events.get(b, {}).values():
            o.comm_stream.wait_event(ev)
        with torch.cuda.stream(o.comm_stream):
            _avg_(self.flat[lo:hi], o.group)

    def finish(self):
        """Reduce the buckets that are still open (they hold parameters that got no gradient in this pass, or whose
        gradient was copied in at the end) and order the current stream after the communication stream."""
        o = self.o
        for b in range(len(o.buckets)):
            if not self.launched[b]:
                self._note(b, None)
                self._launch(b)
        if o.comm_stream is not None:
            torch.cuda.current_stream(o.device).wait_stream(o.comm_stream)
            self.flat.record_stream(o.comm_stream)
        for p, g in self.extra:                       # rare: a parameter used twice inside one pass
            _avg_(g, o.group)
            self.view(p).add_(g)


class BucketedReducer:
    """Gradient buckets over ``params`` in reverse order (the order a backward pass produces them)."""

    def __init__(self, params, bucket_bytes=8 << 20, group=None):
        params = [p for p in params if p.requires_grad]
        self.params = params
        self.group = group
        self.enabled = True
        dev = params[0].device
        self.device = dev
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.slot = {}                # id(param) -> (offset, numel, bucket)
        self.buckets = []             # [lo, hi) element ranges of the flat buffer
        self.bucket_sizes = []
        off = lo = 0
        count = 0
        for p in reversed(params):
            n = p.numel()
            self.slot[id(p)] = (off, n, len(self.buckets))
            off += (n + 3) // 4 * 4   # 16-byte aligned slices
            count += 1
            if (off - lo) * 4 >= bucket_bytes:
                self.buckets.append((lo, off))
                self.bucket_sizes.append(count)
                lo, count = off, 0
        if count:
            self.buckets.append((lo, off))
            self.bucket_sizes.append(count)
        self.total = off

    def begin_pass(self):
        """A pass object for one backward pass, or None while reductions are switched off (no_sync)."""
        return _Pass(self) if self.enabled and _world() > 1 else None


class DataParallel(nn.Module):
    """Drop-in for ``torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])`` on the dualvar_b200
    models (see the module docstring)."""

    def __init__(self, module, device_ids=None, bucket_cap_mb=8, process_group=None):
        super().__init__()
        self.module = module
        self.group = process_group
        self.require_backward_grad_sync = True
        self._cb_queued = False
        if _world() > 1:
            with torch.no_grad():     # DDP's _sync_module_states: everybody starts from rank 0's parameters and buffers
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=0, group=process_group)
        self._reducers = []
        managed = set()
        for m in module.modules():
            if hasattr(m, "encode") and hasattr(m, "program"):           # an engine-run encoder (backbones._Encoder)
                ps = [p for p in m.parameters() if p.requires_grad and id(p) not in managed]
                if ps and ps[0].is_cuda:
                    r = BucketedReducer(ps, bucket_cap_mb << 20, process_group)
                    r._owner = self
                    m._dv_reducer = r
                    self._reducers.append(r)
                    managed.update(id(p) for p in ps)
        self._rest = [p for p in module.parameters() if p.requires_grad and id(p) not in managed]
        self._rest_stream = self._reducers[0].comm_stream if self._reducers else None
        for p in self._rest:
            p.register_post_accumulate_grad_hook(self._on_grad)

    # ---- the gradients autograd produces outside the engine
    def _on_grad(self, _p):
        self.queue_final_callback()

    def queue_final_callback(self):
        """Called from inside a backward pass (parameter hooks, engine.BackboneFunction.backward): reduce the non-engine
        gradients once, when the autograd engine has finished this backward."""
        if not self._cb_queued and self.require_backward_grad_sync and _world() > 1:
            self._cb_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

    def _finalize(self):
        self._cb_queued = False
        ps = [p for p in self._rest if p.grad is not None]
        if not ps:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        _avg_(flat, self.group)
        off = 0
        for p in ps:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n

    @contextlib.contextmanager
    def no_sync(self):
        old = self.require_backward_grad_sync
        self.require_backward_grad_sync = False
        for r in self._reducers:
            r.enabled = False
        try:
            yield
        finally:
            self.require_backward_grad_sync = old
            for r in self._reducers:
                r.enabled = old

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
