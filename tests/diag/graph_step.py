"""GPU diagnostic: is a pretraining step host-bound? The same SimCLR+DualVar step issued eagerly (ctypes calls from Python)
and replayed as ONE CUDA graph (torch.cuda.CUDAGraph over the whole step: ingest, both backbone passes, heads, losses,
backward, fused SGD). Prints ms/step for both and the host issue time of the eager step.

    python tests/diag/graph_step.py [net] [B] [T] [H]
"""
import os, sys, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import models as PM
from dualvar_b200.engine import RawClips
from dualvar_b200.optim import SGD

net = sys.argv[1] if len(sys.argv) > 1 else "s3dg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32
H = int(sys.argv[4]) if len(sys.argv) > 4 else 128
dev = "cuda:0"
torch.manual_seed(0); np.random.seed(0); random.seed(0)
model = PM.SimCLR_TimeSeriesV4(net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                               SimpleNamespace(shufflerank_theta=0.05)).to(dev).train()
opt = SGD([{"params": p} for p in model.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)
frames = torch.rand(B, 3, 3 * T, H, H, device=dev)
perm_static = torch.from_numpy(np.array([np.random.permutation(2) for _ in range(B)], dtype=np.int32)).to(dev)
PM._draw_perms = lambda B_, s, d: perm_static          # the per-step permutation becomes a static input buffer


def step(set_to_none):
    ret = model(RawClips(frames, 3))
    loss = sum(v for k, v in ret.items() if "loss" in k)
    opt.zero_grad(set_to_none=set_to_none)
    loss.backward()
    opt.step()
    return loss


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    t_issue = (time.perf_counter() - t0) / n
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_issue * 1e3


for _ in range(3):
    step(True)
ms, issue = timed(lambda: step(True), 5)
print(f"{net} B={B} {T}x{H}x{H}: eager {ms:.2f} ms/step ({B / ms * 1e3:.1f} samples/s), host issue time {issue:.2f} ms/step", flush=True)
# graph: gradients accumulate into static .grad tensors (zeroed in place), so the optimizer's pointer table is static too
for _ in range(2):
    step(False)
ms2, issue2 = timed(lambda: step(False), 5)
print(f"eager with in-place zero_grad: {ms2:.2f} ms/step, host {issue2:.2f}", flush=True)
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step(False)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        loss = step(False)
    torch.cuda.synchronize()
    l0 = float(loss)
    if os.environ.get("DV_NCU"):      # ncu --profile-from-start off: the kernels of ONE replay, without host gaps
        torch.cuda.profiler.start(); g.replay(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    ms3, issue3 = timed(g.replay, 10)
    print(f"CUDA graph replay: {ms3:.2f} ms/step ({B / ms3 * 1e3:.1f} samples/s), host {issue3:.3f} ms; loss after capture {l0:.4f} "
          f"after replays {float(loss):.4f}", flush=True)
except Exception as e:  # noqa: BLE001
    print("graph capture failed:", type(e).__name__, str(e)[:600], flush=True)
