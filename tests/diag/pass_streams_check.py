"""GPU diagnostic: the two-stream issue of the two backbone passes (engine.BackbonePairFunction) against the sequential
one - same weights, input and permutations: losses, parameter gradients and BN running statistics must agree to the
run-to-run noise of the atomics (compare with a second sequential run)."""
import os, sys, random, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import engine as E, models as PM
from dualvar_b200.engine import RawClips
dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0); np.random.seed(0); random.seed(0)
model = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                               SimpleNamespace(shufflerank_theta=0.05)).to(dev).train()
init = copy.deepcopy(model.state_dict())
frames = torch.rand(B, 3, 48, 112, 112, device=dev)

def run(pair):
    E.PASS_STREAMS = pair
    model.load_state_dict(init)
    np.random.seed(5)
    for p in model.parameters(): p.grad = None
    ret = model(RawClips(frames, 3))
    loss = sum(v for k, v in ret.items() if "loss" in k)
    loss.backward()
    torch.cuda.synchronize()
    return ({k: v.detach().clone() for k, v in ret.items() if "loss" in k},
            {n: p.grad.detach().clone() for n, p in model.named_parameters()},
            {n: b.detach().clone() for n, b in model.named_buffers()})

def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()
seq1, seq2 = run(False), run(False)
for it in range(3):
    pr = run(True)
    noise = sorted(rel(seq2[1][n], seq1[1][n]) for n in seq1[1])
    diff = sorted(rel(pr[1][n], seq1[1][n]) for n in seq1[1])
    worst = max(seq1[1], key=lambda n: rel(pr[1][n], seq1[1][n]))
    bdiff = max(rel(pr[2][n], seq1[2][n]) for n in seq1[2] if seq1[2][n].dtype.is_floating_point)
    nbt = all(int(pr[2][n]) == int(seq1[2][n]) for n in seq1[2] if not seq1[2][n].dtype.is_floating_point)
    print(f"run {it}: loss diff " + " ".join(f"{k.split('_')[0]}={abs(pr[0][k].item()-seq1[0][k].item()):.2e}" for k in seq1[0]) +
          f" | grad rel diff median {diff[len(diff)//2]:.2e} max {diff[-1]:.2e} ({worst}) vs seq-seq noise median {noise[len(noise)//2]:.2e} max {noise[-1]:.2e}"
          f" | running-stat max rel diff {bdiff:.2e} num_batches_tracked equal {nbt}", flush=True)
