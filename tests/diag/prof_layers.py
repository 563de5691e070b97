"""GPU diagnostic: per-kernel / per-layer device time of one pretrain step (CUDA events per C-ABI call)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import _lib, models as PM
from dualvar_b200.engine import RawClips

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net_name = sys.argv[2] if len(sys.argv) > 2 else "r21d"
dev = "cuda:0"
torch.manual_seed(0); np.random.seed(0); random.seed(0)
model = PM.SimCLR_TimeSeriesV4(net_name, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                               SimpleNamespace(shufflerank_theta=0.05)).to(dev).train()
from dualvar_b200.optim import SGD
opt = SGD([{'params': p} for p in model.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16          # frames per clip
H = int(sys.argv[4]) if len(sys.argv) > 4 else 112
frames = torch.rand(B, 3, 3 * T, H, H, device=dev)

def step():
    ret = model(RawClips(frames, 3))
    loss = sum(v for k, v in ret.items() if "loss" in k)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()

for _ in range(2): step()
torch.cuda.synchronize()
timer = _lib.KernelTimer(_lib.exported_symbols(), detail=True)
_lib.set_timer(timer)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record()
_lib.set_timer(None)
summ = timer.summary()
total = e0.elapsed_time(e1)
by_kernel = {}
for k, d in summ.items():
    n = k.split(" ")[0]
    a = by_kernel.setdefault(n, [0, 0.0, 0.0]); a[0] += d["calls"]; a[1] += d["ms"]; a[2] += d["flops"]
print(f"step {total:.2f} ms (with per-call events), B={B} {net_name}")
print("== by kernel ==")
acc = 0
for n, (c, ms, fl) in sorted(by_kernel.items(), key=lambda x: -x[1][1]):
    acc += ms
    print(f"{ms:8.3f} ms {c:5d} calls {fl/ms/1e9 if fl else 0:8.1f} TF/s  {n}")
print(f"sum of timed calls {acc:.2f} ms")
print("== by layer (top 60) ==")
for k, d in sorted(summ.items(), key=lambda x: -x[1]["ms"])[:60]:
    print(f"{d['ms']:8.3f} ms {d['calls']:3d}x {d['flops']/d['ms']/1e9 if d['flops'] else 0:8.1f} TF/s  {k}")
