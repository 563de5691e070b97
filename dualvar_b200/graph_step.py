"""A whole training step as ONE CUDA graph.

An eager step of this package is 500-2300 C-ABI calls issued from Python (ctypes marshalling, one torch allocation per
output, stream bookkeeping). For the Inception-style S3D-G that host work is LONGER than the device work - the GPU idles
between launches (measured on B200, 16 samples of 32x128x128: 76.3 ms eager, 43.6 ms as a graph) - and even for
R(2+1)D the ~700 launch gaps of a step cost 2 % (74.7 -> 73.0 ms at 64 samples). ``GraphedTrainStep`` captures

    ret = model(RawClips(frames, n_views)); loss = loss_fn(ret); zero_grad; loss.backward(); optimizer.step()

once (ingest, every encoder pass, heads, objectives, backward on main + weight-gradient side stream, fused SGD) and
replays it per batch. Everything a step reads from the host becomes a static device buffer refreshed before the replay:

* ``frames``  - the loader batch (fp32 or uint8, pinned host or device memory) is copied into a fixed device tensor;
* the per-sample segment permutations the models draw with ``np.random.permutation`` (model/simclr.py:379-381,
  model/moco.py:544-546) are drawn on the host before every replay, in the reference's order, and copied into the
  fixed tensor the captured ingest kernel reads (``models._perm_source``).

What is baked into a graph and therefore checked before every replay: the batch shape (another shape runs eagerly), the
(lr, momentum, weight_decay) of every param group (a change - MultiStepLR at a milestone - triggers a re-capture).
The first ``warmup`` calls run eagerly (they are real training steps; they also populate every lazily built table) and
gradients live in fixed ``.grad`` tensors that are zeroed in place, so the optimizer's pointer table is static.
Returned tensors are the graph's own output buffers: valid until the next call.

Data-parallel runs: the step is captured per rank with its collectives inside - NCCL all-gathers of the embeddings,
the bucketed gradient all-reduce of ``parallel.DataParallel`` and the cross-replica BatchNorm peer exchanges, which
number their calls from a device-side counter for exactly this purpose (csrc/comm.cu). Every rank must capture and
replay in lockstep (they do: same loop). Under torch's own DistributedDataParallel the step stays eager (its reducer
needs its own warm-up protocol for capture). All eager warm-up steps and the capture run on ONE dedicated stream, so
that per-stream resources (the weight-gradient side stream, the BatchNorm exchange channel) exist before the capture.
"""
import gc
import os

import numpy as np
import torch
import torch.distributed as dist

from . import models as PM
from .engine import RawClips, direct_param_grads
from .pretrain_loop import total_loss


STEP_STREAM_PRIORITY = int(os.environ.get("DV_STEP_STREAM_PRIORITY", "-1"))


class GraphedTrainStep:
    def __init__(self, model, optimizer, n_views=3, loss_fn=total_loss, warmup=3, wrap=None):
        self.model, self.opt, self.n_views, self.loss_fn = model, optimizer, n_views, loss_fn
        self.wrap = wrap if wrap is not None else (lambda fr: RawClips(fr, n_views))
        self.warmup_left = max(1, warmup)      # at least one eager step on the capture stream (see the module docstring)
        self.stream = None
        self.graph = None
        self.frames = None
        self.perm = self.perm_host = None
        self.sig = None
        self.out = None
        target = model.module if hasattr(model, "module") else model
        # graph_safe: the model's forward passes no per-step host value to a kernel by value (MoCo's queue pointer is one)
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.enabled = bool(getattr(target, "graph_safe", False)) and \
            not (distributed and isinstance(model, torch.nn.parallel.DistributedDataParallel))
        self.direct = not distributed and os.environ.get("DV_DIRECT_GRADS", "1") != "0"
        self.n_series = getattr(target, "n_series", 0) or 0
        self.replays = self.eager_steps = self.captures = 0
        self.launches_per_step = 0          # kernel launches recorded into the graph (C-ABI launch counter during capture)

    # ---- host-side random numbers of a step, in the reference's order
    def _draw_perms(self, B):
        if self.n_series <= 0:
            return
        vals = np.array([np.random.permutation(self.n_series) for _ in range(B)], dtype=np.int32)
        self.perm_host.copy_(torch.from_numpy(vals))
        self.perm.copy_(self.perm_host, non_blocking=True)

    def _perm_source(self, B, n_series, device):
        assert self.perm is not None and tuple(self.perm.shape) == (B, n_series)
        return self.perm

    def _lr_signature(self):
        return tuple((g["lr"], g.get("momentum", 0.0), g.get("weight_decay", 0.0)) for g in self.opt.param_groups)

    def _backward(self, loss):
        """loss.backward() with the backbone's parameter gradients added straight into the (zeroed) .grad tensors
        (engine.direct_param_grads: no per-parameter accumulation kernels of autograd); single process only."""
        if self.direct:
            with direct_param_grads(self.model.parameters()):
                loss.backward()
        else:
            loss.backward()

    def _step(self, set_to_none):
        ret = self.model(self.wrap(self.frames))
        loss = self.loss_fn(ret)
        self.opt.zero_grad(set_to_none=set_to_none)
        self._backward(loss)
        self.opt.step()
        out = dict(ret)
        out["loss"] = loss.detach()
        return out

    def _eager(self, frames):
        """A plain step on the caller's tensors (warm-up calls, odd batch shapes, refused models), issued on the
        dedicated stream when graphs are in use."""
        self.eager_steps += 1
        if not self.enabled:
            return self._eager_body(frames)
        cur = torch.cuda.current_stream()
        if self.stream is None:
            # high priority: the weight-gradient side stream (engine._side_stream, default priority) must never hold back
            # a kernel of the critical path - conv / BatchNorm chain - when both have CTAs waiting for an SM
            self.stream = torch.cuda.Stream(priority=STEP_STREAM_PRIORITY)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            out = self._eager_body(frames)
        cur.wait_stream(self.stream)
        return out

    def _eager_body(self, frames):
        ret = self.model(self.wrap(frames))
        loss = self.loss_fn(ret)
        self.opt.zero_grad(set_to_none=False)
        self._backward(loss)
        self.opt.step()
        out = dict(ret)
        out["loss"] = loss.detach()
        return out

    def _capture(self):
        self.graph = None
        self.out = None
        g = torch.cuda.CUDAGraph()
        prev = PM._perm_source
        PM._perm_source = self._perm_source
        from . import _lib
        n0 = _lib.load().dv_launch_count()
        # no cyclic garbage collection while the stream captures: a collection that happens to release CUDA objects of an
        # earlier (e.g. failed and abandoned) graph frees device memory, which invalidates the running capture
        gc_was_on = gc.isenabled()
        try:
            torch.cuda.synchronize()
            gc.collect()
            gc.disable()
            # "relaxed": a cudaFree / cudaMalloc that torch's allocator or a destructor issues while this stream captures
            # (memory of an earlier, released graph going back to the driver) is legal instead of invalidating the capture
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode="relaxed"):
                out = self._step(set_to_none=False)
        except Exception as e:
            raise RuntimeError(
                "CUDA-graph capture of the training step failed. A frequent cause: a tensor WITH autograd history from an "
                "eager step issued on another stream is still referenced (a kept `loss`, a kept output dict) - the "
                "parameters' AccumulateGrad nodes then belong to that stream and autograd synchronises the capturing stream "
                "with it. Detach or drop such tensors before the first call of GraphedTrainStep.") from e
        finally:
            if gc_was_on:
                gc.enable()
            PM._perm_source = prev
        self.launches_per_step = int(_lib.load().dv_launch_count() - n0)
        self.graph, self.out, self.sig = g, out, self._lr_signature()
        # the graph writes into THESE gradient tensors: keep them alive and re-attach them if an eager step in between
        # replaced them (zero_grad(set_to_none=True))
        self._params = [p for grp in self.opt.param_groups for p in grp["params"]]
        self._grads = [p.grad for p in self._params]
        self.captures += 1

    def __call__(self, frames):
        if not self.enabled or not frames.dtype in (torch.float32, torch.uint8):
            return self._eager(frames.to(next(self.model.parameters()).device, non_blocking=True))
        dev = next(self.model.parameters()).device
        if self.warmup_left > 0:
            self.warmup_left -= 1
            return self._eager(frames.to(dev, non_blocking=True))
        if self.frames is None:
            self.frames = torch.empty(frames.shape, dtype=frames.dtype, device=dev)
            if self.n_series > 0:
                self.perm = torch.zeros((frames.shape[0], self.n_series), dtype=torch.int32, device=dev)
                self.perm_host = torch.zeros((frames.shape[0], self.n_series), dtype=torch.int32).pin_memory()
        if tuple(frames.shape) != tuple(self.frames.shape) or frames.dtype != self.frames.dtype:
            return self._eager(frames.to(dev, non_blocking=True))
        if self.graph is None or self.sig != self._lr_signature():
            self._capture()
        for p_, g_ in zip(self._params, self._grads):
            if p_.grad is not g_:
                p_.grad = g_
        self.frames.copy_(frames, non_blocking=True)
        self._draw_perms(frames.shape[0])
        self.graph.replay()
        self.replays += 1
        return self.out

    def release(self):
        """Drop the graph and its private memory pool (activations of a whole step)."""
        self.graph = self.out = None
        self.frames = self.perm = None
        self._params = self._grads = None
