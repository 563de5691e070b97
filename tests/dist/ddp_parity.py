"""Multi-GPU parity (run under torchrun, one rank per GPU): dualvar_b200 SimCLR+DualVar with
SyncBatchNorm + DDP (pretrain.py:244-248) vs the oracle wrapped the same way, same per-rank data.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist/ddp_parity.py
    DV_PRECISION=fp32 ... ddp_parity.py [simclr|moco]     # the fp32 mode against the 1e-4 bar
    ... ddp_parity.py simclr overlap     # the product under dualvar_b200.parallel.DataParallel (bucketed all-reduce inside
                                         # the engine's backward) instead of torch's DDP; its gradients are also compared
                                         # with the product's own torch-DDP gradients (same data): equal up to fp32 sum order
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

FP32 = os.environ.get("DV_PRECISION", "bf16") == "fp32"
# tolerances: bf16 mode 1e-2 (north star), fp32 mode 1e-4; the absolute terms cover losses near zero (a fresh MoCo has
# key encoder == query encoder: the clip loss is a ~1e-3 difference of two ~14 logits)
LOSS_RTOL, LOSS_ATOL, LOGIT_TOL, GRAD_MED = (1e-4, 2e-6, 1e-4, 2e-2) if FP32 else (1e-2, 2e-3, 2e-2, 0.6)


def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
args = SimpleNamespace(shufflerank_theta=0.05)
which = sys.argv[1] if len(sys.argv) > 1 else "simclr"
wrapper = sys.argv[2] if len(sys.argv) > 2 else "ddp"
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
if wrapper == "overlap":
    import copy
    from dualvar_b200.parallel import DataParallel
    prod_twin = copy.deepcopy(prod)                     # the same product model under torch DDP, for a tight gradient check
    twin_ddp = torch.nn.parallel.DistributedDataParallel(prod_twin, device_ids=[local])
    prod_ddp = DataParallel(prod, device_ids=[local])
else:
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
        good = abs(rp[k].item() - rr[k].item()) <= LOSS_RTOL * abs(rr[k].item()) + LOSS_ATOL
    else:
        good = rr[k].shape == rp[k].shape and ((rp[k] - rr[k]).abs().max() / rr[k].abs().max()).item() < LOGIT_TOL
    ok &= bool(good)
    if "loss" in k or not good:
        print(f"[rank {rank}] {k}: ref {rr[k].float().mean().item():.7f} prod {rp[k].float().mean().item():.7f} {'ok' if good else 'MISMATCH'}", flush=True)
lr = sum(v for k, v in rr.items() if "loss" in k); lp = sum(v for k, v in rp.items() if "loss" in k)
lr.backward(); lp.backward()
errs = sorted(rel(pp.grad, pr.grad) for pr, pp in zip(ref.parameters(), prod.parameters()) if pr.grad is not None)
# DDP-averaged gradients must be identical across ranks
g0 = torch.cat([p.grad.flatten() for p in prod.parameters() if p.grad is not None])
gs = [torch.empty_like(g0) for _ in range(world)]
dist.all_gather(gs, g0)
same = all(torch.equal(gs[0], g) for g in gs)
print(f"[rank {rank}] median grad rel err vs oracle-DDP {errs[len(errs)//2]:.3e}; grads identical across ranks: {same}", flush=True)
ok &= same and errs[len(errs) // 2] < GRAD_MED
if wrapper == "overlap":
    torch.manual_seed(50)
    np.random.seed(10 * rank + 1); rt = twin_ddp(x)
    sum(v for k, v in rt.items() if "loss" in k).backward()
    # same kernels, same data: the two reductions differ by fp32 summation order (and the atomics order of a second run)
    worst = max(rel(pp.grad, pt.grad) for pp, pt in zip(prod.parameters(), prod_twin.parameters()) if pt.grad is not None)
    missing = [n for (n, pp), pt in zip(prod.named_parameters(), prod_twin.parameters()) if (pp.grad is None) != (pt.grad is None)]
    print(f"[rank {rank}] overlapped reducer vs torch DDP on the product: worst per-tensor rel diff {worst:.3e}, missing {missing}", flush=True)
    ok &= worst < 2e-2 and not missing
t = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DDP_PARITY", which, "fp32 mode" if FP32 else "bf16 mode", "PASS" if t.item() == 1.0 else "FAIL",
          "(dualvar_b200.parallel.DataParallel)" if wrapper == "overlap" else "", flush=True)
dist.destroy_process_group()
