"""Target for the per-layer ncu capture of one bench step (DRAM bytes + duration of EVERY launch, named by layer).

    DV_NVTX=1 ncu --nvtx --print-nvtx-rename kernel --profile-from-start off --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
        --log-file gpurun_out/ncu_step.csv python tests/diag/ncu_step.py
    python tests/diag/ncu_traffic.py gpurun_out/ncu_step.csv profiles/r02_ncu_traffic.json

With DV_NVTX=1 every C-ABI call sits in an NVTX range named "<entry point>_<geometry>" (dualvar_b200/_lib.py), and
--print-nvtx-rename kernel makes ncu report that name instead of the CUDA kernel's, so launches map to the layers of
bench.py's roofline.layers table. The step is the bench step (r21d SimCLR+DualVar, B samples of 3 x 16x112x112,
SGD) on one stream; warm-up steps run outside the profiled region (cudaProfilerStart/Stop).
"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import engine as E, models as PM
from dualvar_b200.engine import RawClips
from dualvar_b200.optim import SGD

dev = "cuda:0"
B = int(os.environ.get("B", "64"))
E.WGRAD_SIDE_STREAM = False          # one stream: ncu serialises launches anyway
torch.manual_seed(0); np.random.seed(0); random.seed(0)
model = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                               SimpleNamespace(shufflerank_theta=0.05)).to(dev).train()
opt = SGD([{"params": p} for p in model.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)
frames = torch.rand(B, 3, 48, 112, 112, device=dev)


def step():
    ret = model(RawClips(frames, 3))
    loss = sum(v for k, v in ret.items() if "loss" in k)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu_step ok, loss", float(loss.detach()))
