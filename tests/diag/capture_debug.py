"""GPU diagnostic: which call invalidates a CUDA-graph capture of a MoCo+DualVar step? Every C-ABI call is followed by
cudaStreamIsCapturing on the capture stream; the first call after which the status is 'invalidated' is printed."""
import os, sys, ctypes, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import _lib, engine as E, models as PM
from dualvar_b200.engine import RawClips
from dualvar_b200.optim import SGD
from dualvar_b200.graph_step import GraphedTrainStep
dev = "cuda:0"
B = int(os.environ.get("B", "64"))
K = int(os.environ.get("K", "16384"))
rt = ctypes.CDLL("libcudart.so.12")
state = {"bad": False, "n": 0}
orig_call = _lib.call


def status():
    st = ctypes.c_int(0)
    rt.cudaStreamIsCapturing(ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(0)), ctypes.byref(st))
    return st.value


def call(name, *args):
    before = status()
    try:
        orig_call(name, *args)
    finally:
        after = status()
        state["n"] += 1
        if after == 2 and not state["bad"]:
            state["bad"] = True
            print(f"capture INVALIDATED at C-ABI call #{state['n']} {name} (status before {before})", flush=True)
            print("args:", [_lib._fmt(a) for a in args], flush=True)
            traceback.print_stack(limit=12)


_lib.call = call
E.call = call
import dualvar_b200.objectives as O, dualvar_b200.models as M2, dualvar_b200.optim as OP
for mod in (O, M2, OP):
    if hasattr(mod, "call"):
        mod.call = call
a = SimpleNamespace(shufflerank_theta=0.05)
torch.manual_seed(0); np.random.seed(0)
m = PM.MoCo_TimeSeriesV4("r21d", 128, K, 0.999, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", a).to(dev).train()
T = int(os.environ.get("T", "48")); HW = int(os.environ.get("HW", "112"))
frames = torch.rand(B, 3, T, HW, HW, device=dev)
opt = SGD([{"params": p} for p in m.parameters() if p.requires_grad], lr=0.003, weight_decay=1e-4, momentum=0.9)
if os.environ.get("EAGER_FIRST"):
    for i in range(int(os.environ["EAGER_FIRST"])):
        ret = m(RawClips(frames, 3))
        loss = sum(v for k, v in ret.items() if "loss" in k)
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    loss = float(loss.detach()); del ret
    print("eager steps done", loss, flush=True)
gs = GraphedTrainStep(m, opt, n_views=3, warmup=1)
try:
    for i in range(4):
        out = gs(frames)
        print("step", i, float(out["loss"]), "captures", gs.captures, flush=True)
except Exception as e:
    print("FAILED:", type(e).__name__, str(e)[:200])
