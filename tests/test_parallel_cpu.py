"""CPU, world_size 2 over gloo: the host logic of dualvar_b200.parallel - bucket layout in reverse parameter order,
buckets reduced as soon as their last gradient is reported, late / unused / twice-used parameters, averaging, and the
DataParallel wrapper's end-of-backward reduction of the gradients autograd produces outside the engine."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


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
        from dualvar_b200 import parallel as DP
        torch.manual_seed(0)
        params = [nn.Parameter(torch.zeros(s)) for s in ((64, 32), (32,), (128, 64), (7,), (300, 100), (5,))]
        red = DP.BucketedReducer(params, bucket_bytes=16 << 10)
        out = {"n_buckets": len(red.buckets), "first_slot_is_last_param": red.slot[id(params[-1])][0] == 0,
               "aligned": all(off % 4 == 0 for off, _, _ in red.slot.values())}
        # a pass: gradients written into the views in backward order; rank r contributes (r + 1) * k for parameter k
        rp = red.begin_pass()
        launched_during = []
        for k in reversed(range(len(params))):
            if k == 1:
                continue                                    # parameter 1 gets no gradient in this pass (unused)
            rp.view(params[k]).fill_(float((rank + 1) * (k + 1)))
            rp.ready(params[k])
            launched_during.append(sum(rp.launched))
        out["overlap"] = launched_during[0] <= launched_during[-1] and max(launched_during) >= 1
        rp.extra.append((params[0], torch.full_like(params[0], float(rank + 1))))   # a second contribution, late
        rp.finish()
        mean_rank = (1 + world) / 2.0                       # average of (r + 1)
        ok = True
        for k, p in enumerate(params):
            want = 0.0 if k == 1 else mean_rank * (k + 1) + (mean_rank if k == 0 else 0.0)
            ok &= bool(torch.allclose(rp.view(p), torch.full_like(p, want)))
        out["averaged"] = ok
        out["all_launched"] = all(rp.launched)
        # the wrapper: gradients from plain autograd are averaged by the end-of-backward callback; no_sync leaves them
        torch.manual_seed(1)
        net = nn.Sequential(nn.Linear(8, 8), nn.ReLU(), nn.Linear(8, 3))
        with torch.no_grad():
            for p in net.parameters():
                p.add_(rank)                                # ranks start different: the wrapper must broadcast rank 0's
        wrapped = DP.DataParallel(net)
        same = [p.detach().clone() for p in net.parameters()]
        gathered = [torch.zeros_like(same[0]) for _ in range(world)]
        dist.all_gather(gathered, same[0])
        out["broadcast"] = all(torch.equal(gathered[0], g) for g in gathered)
        x = torch.randn(4, 8, generator=torch.Generator().manual_seed(10 + rank))
        wrapped(x).pow(2).sum().backward()
        g_sync = [p.grad.clone() for p in net.parameters()]
        # reference: all ranks' local gradients averaged by hand
        ref_net = nn.Sequential(nn.Linear(8, 8), nn.ReLU(), nn.Linear(8, 3))
        ref_net.load_state_dict(net.state_dict())
        ref_net(x).pow(2).sum().backward()
        ok = True
        for p, g in zip(ref_net.parameters(), g_sync):
            t = p.grad.clone()
            dist.all_reduce(t)
            ok &= bool(torch.allclose(t / world, g, rtol=1e-6, atol=1e-7))
        out["wrapper_average"] = ok
        net.zero_grad()
        with wrapped.no_sync():
            wrapped(x).pow(2).sum().backward()
        out["no_sync_local"] = all(torch.allclose(p.grad, q_.grad) for p, q_ in zip(net.parameters(), ref_net.parameters()))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_bucketed_reducer_and_wrapper_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in results.items():
        assert out.pop("n_buckets") >= 3
        assert all(out.values()), (rank, out)
