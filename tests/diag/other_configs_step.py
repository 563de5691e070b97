import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import models as PM
from dualvar_b200.optim import SGD
dev = "cuda:0"
for net, B, T, H in (("s3dg", 16, 32, 128), ("r3d", 32, 16, 112), ("c3d", 16, 16, 112), ("r2d3d18", 32, 16, 112)):
    torch.manual_seed(0); np.random.seed(0); random.seed(0)
    model = PM.SimCLR_TimeSeriesV4(net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", SimpleNamespace(shufflerank_theta=0.05)).to(dev).train()
    opt = SGD([{'params': p} for p in model.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)
    x = torch.randn(B, 3, 3, T, H, H, device=dev)
    def step():
        ret = model(x); loss = sum(v for k, v in ret.items() if "loss" in k)
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); return loss
    for _ in range(2): l = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): l = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{net}: B={B} {T}x{H}x{H}  {ms:.1f} ms/step  {B/ms*1e3:.1f} samples/s  loss {float(l):.4f}", flush=True)
    del model, opt, x; torch.cuda.empty_cache()
