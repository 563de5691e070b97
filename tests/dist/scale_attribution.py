"""Multi-GPU diagnostic (torchrun, one rank per GPU): where does the weak-scaling loss of the bench step go?

Times the bench step (r21d SimCLR+DualVar, 64 samples per GPU, SyncBatchNorm, resident inputs; replayed as one CUDA graph
per rank like bench.py does - GRAPH=0 issues it eagerly, which hides most of the differences behind host time) in variants:
  full/overlap   dualvar_b200.parallel.DataParallel (bucketed gradient all-reduce inside the backward) - the default
  full/torch     torch DistributedDataParallel (all-reduce after the one-node backbone backward)
  no_grad_sync   gradient all-reduce skipped (no_sync) - what the reduction costs end to end
  bn_local       cross-replica BatchNorm exchanges skipped (statistics stay per rank: WRONG numerics, diagnostic only) -
                 what the ~400 statistic rendezvous per step cost
  local          both skipped = N independent single-GPU steps (the 1-GPU step time on this box under N-GPU load)
Prints one line per variant (max over ranks, CUDA events).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/dist/scale_attribution.py
"""
import os, sys, random, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch, torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from dualvar_b200 import engine as E, models as PM
from dualvar_b200.engine import RawClips
from dualvar_b200.optim import SGD
from dualvar_b200.parallel import DataParallel
from dualvar_b200.graph_step import GraphedTrainStep
GRAPH = os.environ.get("GRAPH", "1") != "0"

B = int(os.environ.get("B", "64"))
STEPS = int(os.environ.get("STEPS", "8"))
frames = torch.rand(B, 3, 48, 112, 112, device=dev, generator=torch.Generator(device=dev).manual_seed(1234 + rank))
real_is_sync = E._is_sync


def build(wrapper):
    torch.manual_seed(0); np.random.seed(0); random.seed(0)
    m = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, True, True, 2, 64, 0.07, 0.07, "clip-sr-tc", SimpleNamespace(shufflerank_theta=0.05))
    m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m).to(dev).train()
    w = DataParallel(m, device_ids=[local]) if wrapper == "overlap" else torch.nn.parallel.DistributedDataParallel(m, device_ids=[local])
    opt = SGD([{"params": p} for p in m.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)
    return w, opt


def run(tag, wrapper, grad_sync=True, bn_sync=True):
    E._is_sync = real_is_sync if bn_sync else (lambda bn: False)
    model, opt = build(wrapper)

    graphed = GraphedTrainStep(model, opt, n_views=3, warmup=1) if GRAPH and wrapper == "overlap" else None

    def step():
        ctx = contextlib.nullcontext() if grad_sync else model.no_sync()
        with ctx:
            if graphed is not None:
                graphed(frames)
                return
            ret = model(RawClips(frames, 3))
            loss = sum(v for k, v in ret.items() if "loss" in k)
            opt.zero_grad(set_to_none=True)
            loss.backward()
        opt.step()
    for _ in range(4):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / STEPS], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{tag:14s} {float(t):8.2f} ms/step  {B * world / float(t) * 1e3:9.1f} samples/s  ({world} GPUs)", flush=True)
    if graphed is not None:
        graphed.release()
    del model, opt, graphed
    torch.cuda.empty_cache()
    E._is_sync = real_is_sync


run("full/overlap", "overlap")
run("full/torch", "torch")
run("no_grad_sync", "overlap", grad_sync=False)
run("bn_local", "overlap", bn_sync=False)
run("local", "overlap", grad_sync=False, bn_sync=False)
dist.destroy_process_group()
