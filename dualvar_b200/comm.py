"""Latency-bound collectives of the data-parallel path on NVLink peer memory.

``small_allreduce_(t)`` sums a small fp64 vector (the per-layer statistics of cross-replica BatchNorm: 2*Cp doubles in
forward, 2*Cp in backward, ~400 vectors per pretraining step, each on the critical path) across the ranks of one node
with one kernel launch (csrc/comm.cu) instead of one NCCL call (~20 us each). The symmetric buffer comes from
``torch.distributed._symmetric_memory`` (plumbing: allocation + peer mapping); the exchange itself is our kernel.

Reference: nn.SyncBatchNorm's per-layer all_gather / all_reduce after convert_sync_batchnorm (pretrain.py:244).
DV_SMALL_ALLREDUCE=0 keeps NCCL. DV_PEER_TIMEOUT_S (default 600, NCCL's watchdog default) bounds how long a rank
waits for a peer; a time-out or ranks with different per-rank batch sizes are reported through ``PeerExchangeError``
at the next exchange instead of trapping the device.
"""
import ctypes
import os
import warnings

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr, stream_ptr

_ENABLED = os.environ.get("DV_SMALL_ALLREDUCE", "1") != "0"
_MAX_ELEMS = 4095  # one element of every slot carries the per-rank count tag
_state = None      # None = not tried, False = unavailable (NCCL is used), else {stream id: _PeerReduce}


class PeerExchangeError(RuntimeError):
    pass


def check_status():
    """Raise if an earlier peer exchange timed out or saw unequal per-rank BatchNorm counts (host-mapped status word
    written by the kernels; reading it costs no synchronisation)."""
    peer = ctypes.c_int(0)
    seq = ctypes.c_int64(0)
    code = _lib.load().dv_comm_status(ctypes.byref(peer), ctypes.byref(seq), 1)
    if code == 1:
        raise PeerExchangeError(f"rank {dist.get_rank()}: peer exchange {seq.value} timed out waiting for rank {peer.value} "
                                "(DV_PEER_TIMEOUT_S)")
    if code == 2:
        raise PeerExchangeError(f"rank {dist.get_rank()}: rank {peer.value} runs a different per-rank batch size in exchange "
                                f"{seq.value}; cross-replica BatchNorm on the peer path needs equal batches on all ranks "
                                "(use drop_last / DistributedSampler padding, or DV_SMALL_ALLREDUCE=0)")


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

    def next_call(self):
        """(peer_buffers, rank, world, seq) of the next exchange: the trailing arguments of every dv_*_sync entry.
        seq = 0: the kernels number the calls themselves (device-side counter in the symmetric buffer), so a captured
        step replays correctly; self.seq only counts for diagnostics."""
        check_status()
        self.seq += 1
        return self.ptrs, self.rank, self.world, 0

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
        _lib.load().dv_comm_set_timeout(ctypes.c_double(float(os.environ.get("DV_PEER_TIMEOUT_S", "600"))))
    # one exchange channel (symmetric buffer + call counter) per issuing stream: calls of one channel execute in
    # order on every rank, which is what makes the two-parity slot reuse safe; concurrent backbone passes
    # (engine.BackbonePairFunction) therefore must not share a channel
    key = torch.cuda.current_stream(device).cuda_stream
    ch = _state.get(key)
    if ch is None:
        err = None
        try:
            ch = _PeerReduce(device)
        except Exception as e:  # noqa: BLE001 - peer mapping is an optimisation; NCCL carries the same sum
            err, ch = e, None
        # the decision is collective: if the mapping failed on ANY rank, every rank falls back to NCCL (a rank that
        # issued dist.all_reduce while the others spin on peer flags would hang both)
        ok = torch.tensor([0 if ch is None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            warnings.warn(f"dualvar_b200: NVLink peer all-reduce unavailable on some rank ({err!r} here); "
                          "using NCCL for BatchNorm statistics on all ranks")
            _state = False
            return None
        torch.cuda.synchronize(device)
        dist.barrier()          # every rank's flags are zero before anybody signals
        _state[key] = ch
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
