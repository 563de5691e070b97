"""GPU diagnostic: BASELINE configs 3 and 5 - one R(2+1)D MoCo+DualVar step (K=16384, m=0.999) and the retrieval maths on
UCF101-shaped feature matrices (3783 x 9537 x 512)."""
import os, sys, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import models as PM, retrieval as R
from dualvar_b200.optim import SGD
from dualvar_b200.engine import RawClips
dev = "cuda:0"
torch.manual_seed(0); np.random.seed(0); random.seed(0)
B = 64
model = PM.MoCo_TimeSeriesV4("r21d", 128, 16384, 0.999, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                             SimpleNamespace(shufflerank_theta=0.05)).to(dev).train()
opt = SGD([{'params': p} for p in model.parameters() if p.requires_grad], lr=0.003, weight_decay=1e-4, momentum=0.9)
frames = torch.rand(B, 3, 48, 112, 112, device=dev)
def step():
    ret = model(RawClips(frames, 3)); loss = sum(v for k, v in ret.items() if "loss" in k)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); return loss
for _ in range(2): l = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"MoCo+DualVar r21d K=16384: B={B}  {ms:.1f} ms/step  {B/ms*1e3:.1f} samples/s  loss {float(l.detach()):.4f}", flush=True)
del model, opt
test = torch.randn(3783, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
train = torch.randn(9537, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(8))
for _ in range(2): out = R.retrieval_topk(test, train)
torch.cuda.synchronize()
e0.record(); out = R.retrieval_topk(test, train); e1.record(); torch.cuda.synchronize()
print(f"retrieval 3783 x 9537 x 512 (centre, normalise, fp64 similarity, top-1/5/10/20/50): {e0.elapsed_time(e1):.2f} ms", flush=True)
t0 = time.perf_counter()
tc, trc = test.cpu(), train.cpu()
tm = trc.mean(0, keepdim=True); a = torch.nn.functional.normalize(tc - tm, dim=1); b = torch.nn.functional.normalize(trc - tm, dim=1)
sim = a @ b.t(); idx = [torch.topk(sim, k, dim=1)[1] for k in (1, 5, 10, 20, 50)]
print(f"same maths with torch on the host CPU ({torch.get_num_threads()} threads): {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
