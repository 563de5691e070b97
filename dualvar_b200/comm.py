"""Latency-bound collectives of the data-parallel path on NVLink peer memory.

``small_allreduce_(t)`` sums a small fp64 vector (the per-layer statistics of cross-replica BatchNorm: 2*Cp doubles in
forward, 2*Cp in backward, ~400 vectors per pretraining step, each on the critical path) across the ranks of one node
with one kernel launch (csrc/comm.cu) instead of one NCCL call (~20 us each). The symmetric buffer comes from
``torch.distributed._symmetric_memory`` (plumbing: allocation + peer mapping); the exchange itself is our kernel.

Reference: nn.SyncBatchNorm's per-layer all_gather / all_reduce after convert_sync_batchnorm (pretrain.py:244).
DV_SMALL_ALLREDUCE=0 keeps NCCL.
"""
import ctypes
import os
import warnings

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr, stream_ptr

_ENABLED = os.environ.get("DV_SMALL_ALLREDUCE", "1") != "0"
_MAX_ELEMS = 4096
_state = None      # None = not tried, False = unavailable (NCCL is used), else {stream id: _PeerReduce}


class _PeerReduce:
    def __init__(self, device):
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        if self.world > 8:
            raise RuntimeError("single node, at most 8 ranks")
        nbytes = int(_lib.load().dv_allreduce_small_buffer_bytes())
        self.buf = symm.empty(nbytes // 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != self.world or int(ptrs[self.rank]) != self.buf.data_ptr():
            raise RuntimeError("unexpected symmetric-memory mapping")
        self.ptrs = (ctypes.c_int64 * 8)(*[int(p) for p in ptrs], *([0] * (8 - self.world)))
        self.seq = 0
        torch.cuda.synchronize(device)
        dist.barrier()          # every rank's flags are zero before anybody signals

    def next_call(self):
        """(peer_buffers, rank, world, seq) of the next exchange: the trailing arguments of every dv_*_sync entry."""
        self.seq += 1
        return self.ptrs, self.rank, self.world, self.seq

    def __call__(self, t):
        _lib.call("dv_allreduce_small_f64", ptr(t), t.numel(), *self.next_call(), stream_ptr())


def peer_state(device, n):
    """The peer-memory exchange object if it is usable for vectors of n doubles on this process group, else None
    (NCCL is then used). Set up lazily on first use - collectively, so all ranks must reach it together."""
    global _state
    if _state is False or n > _MAX_ELEMS:
        return None
    if _state is None:
        _state = {}
        if not (_ENABLED and dist.get_backend() == "nccl"):
            _state = False
            return None
    # one exchange channel (symmetric buffer + call counter) per issuing stream: calls of one channel execute in
    # order on every rank, which is what makes the two-parity slot reuse safe; concurrent backbone passes
    # (engine.BackbonePairFunction) therefore must not share a channel
    key = torch.cuda.current_stream(device).cuda_stream
    ch = _state.get(key)
    if ch is None:
        try:
            ch = _state[key] = _PeerReduce(device)
        except Exception as e:  # noqa: BLE001 - peer mapping is an optimisation; NCCL carries the same sum
            warnings.warn(f"dualvar_b200: NVLink peer all-reduce unavailable ({e!r}); using NCCL for BatchNorm statistics")
            _state = False
            return None
    return ch


def small_allreduce_(t):
    """In-place sum of a contiguous fp64 CUDA vector over the default process group."""
    assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()
    peer = peer_state(t.device, t.numel())
    if peer is not None:
        peer(t)
    else:
        dist.all_reduce(t)
    return t


def reset():
    """Forget the peer mapping (tests; call on all ranks before destroying the process group)."""
    global _state
    _state = None
