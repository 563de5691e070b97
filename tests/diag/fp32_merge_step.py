"""fp32 mode step (R(2+1)D-18 SimCLR+DualVar, 16x112x112): merged plane-product launches (one per convolution) against
one launch per plane product, interleaved in one process, plus the per-kernel split of the merged step."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import engine as E, models as PM
from dualvar_b200.optim import SGD
dev = "cuda:0"
B = int(os.environ.get("B", "16"))
PLANES = int(os.environ.get("PLANES", "3"))
a = ("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", SimpleNamespace(shufflerank_theta=0.05))
torch.manual_seed(0); np.random.seed(0); random.seed(0)
E.set_precision("fp32", PLANES)
m = PM.SimCLR_TimeSeriesV4(*a).to(dev).train()
opt = SGD([{"params": p} for p in m.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)
x = torch.randn(B, 3, 3, 16, 112, 112, device=dev)


def step():
    np.random.seed(7)
    ret = m(x); loss = sum(v for k, v in ret.items() if "loss" in k)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); return loss.detach()


def timed(n=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {0: [], 1: []}
for rnd in range(4):
    for merge in (1, 0):
        E.F32_MERGE = bool(merge)
        if rnd == 0:
            step(); step()
        res[merge].append(timed())
for merge in (1, 0):
    ms = min(res[merge])
    print(f"fp32 mode ({PLANES} planes) B={B} {'merged' if merge else 'per-product'}: {ms:.1f} ms/step "
          f"{B / ms * 1e3:.1f} samples/s  (rounds {', '.join(f'{v:.1f}' for v in res[merge])})", flush=True)

E.F32_MERGE = True
from torch.profiler import profile, ProfilerActivity
step()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows)
print(f"merged step, device time by kernel (total {tot / 1e3:.1f} ms):")
for r in rows[:22]:
    print(f"  {r.device_time_total / 1e3:8.2f} ms {r.count:5d}x  {r.key[:110]}")
