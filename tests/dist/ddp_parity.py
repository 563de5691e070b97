"""Multi-GPU parity (run under torchrun, one rank per GPU): dualvar_b200 SimCLR+DualVar with
SyncBatchNorm + DDP (pretrain.py:244-248) vs the oracle wrapped the same way, same per-rank data.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist/ddp_parity.py
"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch, torch.distributed as dist

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from oracle import models as OM
from dualvar_b200 import models as PM

def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
args = SimpleNamespace(shufflerank_theta=0.05)
which = sys.argv[1] if len(sys.argv) > 1 else "simclr"
torch.manual_seed(0); np.random.seed(0); random.seed(0)
if which == "simclr":
    ref = OM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, True, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    prod = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, True, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
else:
    ref = OM.MoCo_TimeSeriesV4("r21d", 128, 64, 0.9, 0.07, True, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    prod = PM.MoCo_TimeSeriesV4("r21d", 128, 64, 0.9, 0.07, True, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
prod.load_state_dict(ref.state_dict())
ref = torch.nn.SyncBatchNorm.convert_sync_batchnorm(ref).to(dev).train()
prod = torch.nn.SyncBatchNorm.convert_sync_batchnorm(prod).to(dev).train()
ref_ddp = torch.nn.parallel.DistributedDataParallel(ref, device_ids=[local])
prod_ddp = torch.nn.parallel.DistributedDataParallel(prod, device_ids=[local])
x = torch.randn(4, 3, 3, 8, 64, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
ok = True
torch.manual_seed(50)      # MoCo shuffle-BN draws torch.randperm on every rank (rank 0's is broadcast)
np.random.seed(10 * rank + 1); rr = ref_ddp(x)
torch.manual_seed(50)
np.random.seed(10 * rank + 1); rp = prod_ddp(x)
for k in rr:
    if "labels" in k:
        good = torch.equal(rr[k], rp[k])
    elif "loss" in k:
        good = abs(rp[k].item() - rr[k].item()) <= 1e-2 * abs(rr[k].item()) + 2e-3
    else:
        good = rr[k].shape == rp[k].shape and ((rp[k] - rr[k]).abs().max() / rr[k].abs().max()).item() < 2e-2
    ok &= bool(good)
    if "loss" in k or not good:
        print(f"[rank {rank}] {k}: ref {rr[k].float().mean().item():.5f} prod {rp[k].float().mean().item():.5f} {'ok' if good else 'MISMATCH'}", flush=True)
lr = sum(v for k, v in rr.items() if "loss" in k); lp = sum(v for k, v in rp.items() if "loss" in k)
lr.backward(); lp.backward()
errs = sorted(rel(pp.grad, pr.grad) for pr, pp in zip(ref.parameters(), prod.parameters()) if pr.grad is not None)
# DDP-averaged gradients must be identical across ranks
g0 = torch.cat([p.grad.flatten() for p in prod.parameters() if p.grad is not None])
gs = [torch.empty_like(g0) for _ in range(world)]
dist.all_gather(gs, g0)
same = all(torch.equal(gs[0], g) for g in gs)
print(f"[rank {rank}] median grad rel err vs oracle-DDP {errs[len(errs)//2]:.3e}; grads identical across ranks: {same}", flush=True)
ok &= same and errs[len(errs) // 2] < 0.6
t = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DDP_PARITY", which, "PASS" if t.item() == 1.0 else "FAIL", flush=True)
dist.destroy_process_group()
