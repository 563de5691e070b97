"""CPU, world_size 2 over gloo: the multi-rank semantics of the objectives that the GPU path must
reproduce: all-gathered negatives, local-slice gradient of GatherLayer, per-rank tc rows.
Also pins the column-order arithmetic the CUDA row kernel uses against the oracle's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import objectives as OO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(7)
        n, d, s, e = 3, 16, 2, 8
        full = F.normalize(torch.randn(world * n, 2, d), dim=-1)
        full_s = F.normalize(torch.randn(world * n, 2, s, e), dim=-1)
        # single-process reference on the concatenated batch
        f1 = full.clone().requires_grad_(True)
        _, _, loss1 = OO.nt_xent(f1, 0.07, distributed=False)
        loss1.backward()
        s1 = full_s.clone().requires_grad_(True)
        logits_t1, _, tc1 = OO.tc_loss(s1, 0.07, distributed=False)
        # distributed
        fl = full[rank * n:(rank + 1) * n].clone().requires_grad_(True)
        logits, labels, loss = OO.nt_xent(fl, 0.07, distributed=True)
        loss.backward()
        sl = full_s[rank * n:(rank + 1) * n].clone().requires_grad_(True)
        logits_t, _, tc = OO.tc_loss(sl, 0.07, distributed=True)
        tc.backward()
        tcs = [torch.zeros(()) for _ in range(world)]
        dist.all_gather(tcs, tc.detach())
        out = {
            "clip_equal": torch.allclose(loss, loss1, rtol=1e-5, atol=1e-6),
            "clip_shape": tuple(logits.shape) == (2 * world * n, 2 * world * n - 1),
            "clip_grad_slice": torch.allclose(fl.grad, f1.grad[rank * n:(rank + 1) * n], rtol=1e-4, atol=1e-6),
            "tc_shape": tuple(logits_t.shape) == (2 * n, 2 * world * n - 1),
            "tc_rank_mean": torch.allclose(torch.stack(tcs).mean(), tc1, rtol=1e-5, atol=1e-6),
            "tc_rows": torch.allclose(
                logits_t, torch.cat([logits_t1[rank * n:(rank + 1) * n],
                                     logits_t1[world * n + rank * n:world * n + (rank + 1) * n]]), rtol=1e-5, atol=1e-5),
        }
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_oracle_distributed_objectives_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in results:
        for k, v in out.items():
            assert v, (rank, k)


@pytest.mark.parametrize("n,N,rank,local", [(4, 4, 0, False), (3, 6, 1, True), (2, 8, 3, True), (5, 5, 0, True)])
def test_cuda_row_kernel_column_order_matches_oracle(n, N, rank, local):
    """contrast_rows_kernel places column c of row r at logits index
    j = 0 if c == pos else 1 + c - [self < c] - [pos < c]; check it equals the oracle's gather order."""
    from dualvar_b200.objectives import _row_indices
    rows, self_col, pos_col = _row_indices(n, N, rank, local, torch.device("cpu"))
    want = OO._positive_first_columns(rows, 2 * N, N)
    for i in range(rows.numel()):
        s, p = int(self_col[i]), int(pos_col[i])
        got = [None] * (2 * N - 1)
        for c in range(2 * N):
            if c == s:
                continue
            j = 0 if c == p else 1 + c - (1 if c > s else 0) - (1 if c > p else 0)
            got[j] = c
        assert got == want[i].tolist()


def _shuffle_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dualvar_b200.models import shuffle_bn_plan, exchange_clips, concat_all_gather
        ok = {}
        for n_local, seed in ((4, 0), (5, 1), (1, 2), (8, 3)):
            torch.manual_seed(seed)
            idx = torch.randperm(n_local * world)
            dist.broadcast(idx, src=0)                      # rank 0's permutation, as model/moco.py:371-374
            plan = shuffle_bn_plan(idx, world, rank, n_local)
            # clip j of the gathered batch carries the value j in every element
            mine = torch.arange(rank * n_local, (rank + 1) * n_local, dtype=torch.float32).view(-1, 1, 1).expand(-1, 3, 2).contiguous()
            got = exchange_clips(mine, plan)
            want_set = sorted(idx.view(world, n_local)[rank].tolist())     # the reference's BatchNorm group of this rank
            ok[f"group_{n_local}"] = sorted(got[:, 0, 0].long().tolist()) == want_set and bool((got == got[:, :1, :1]).all())
            ok[f"order_{n_local}"] = got[:, 0, 0].long().tolist() == plan["effective"][rank].tolist()
            # "keys" = the clip id the encoder saw in each slot; un-shuffled they must be this rank's own clips in order
            keys = got[:, 0, :1].contiguous()
            back = concat_all_gather(keys)[plan["unshuffle_rows"]]
            ok[f"unshuffle_{n_local}"] = back[:, 0].long().tolist() == list(range(rank * n_local, (rank + 1) * n_local))
            ok[f"counts_{n_local}"] = sum(plan["send_counts"]) == n_local and sum(plan["recv_counts"]) == n_local
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_moco_shuffle_bn_all_to_all_plan(world):
    """MoCo's shuffle-BN as ONE all-to-all (models.shuffle_bn_plan / exchange_clips; model/moco.py:357-402): every rank ends
    up with exactly the reference's BatchNorm group, and the un-shuffle returns every key to its own rank and slot."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shuffle_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in results:
        for k, v in out.items():
            assert v, (rank, k)
