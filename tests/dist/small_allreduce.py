"""Multi-GPU check (torchrun, one rank per GPU): dualvar_b200.comm.small_allreduce_ vs NCCL all_reduce, bit for bit
(both sum in a fixed order only in our kernel, so the comparison is against an fp64 gather + ordered sum), plus latency.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/dist/small_allreduce.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from dualvar_b200 import comm

ok = True
gen = torch.Generator(device=dev).manual_seed(7 + rank)
for it in range(300):
    n = [16, 128, 288, 1856, 4096][it % 5]
    x = torch.randn(n, device=dev, dtype=torch.float64, generator=gen)
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    want = torch.zeros_like(x)
    for p in parts:
        want += p
    got = comm.small_allreduce_(x.clone())
    ok &= bool(torch.equal(got, want))
assert comm._state is not False, "peer all-reduce was not used"
# latency
x = torch.randn(288, device=dev, dtype=torch.float64)
for name, fn in [("peer", lambda: comm.small_allreduce_(x)), ("nccl", lambda: dist.all_reduce(x))]:
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per 288-double all-reduce (back to back, device time)", flush=True)
t = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SMALL_ALLREDUCE", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
comm.reset()
dist.destroy_process_group()
