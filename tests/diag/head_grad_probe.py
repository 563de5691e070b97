"""GPU diagnostic: where does the head-gradient difference at the bench geometry (4 samples) come from?
Per loss term: gradient of the projection heads and of the pooled features for product / rounding oracle vs fp32 oracle."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
from bf16_emulation import emulate_bf16, round_conv_weights, round_input
from dualvar_b200 import models as PM
from oracle import models as OM
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
ARGS = SimpleNamespace(shufflerank_theta=0.05)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0); np.random.seed(0); random.seed(0)
ref = OM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS).to(dev).train()
round_conv_weights(ref)
emu = emulate_bf16(ref)
prod = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS)
prod.load_state_dict(ref.state_dict()); prod = prod.to(dev).train()
x = round_input(torch.randn(B, 3, 3, 16, 112, 112, device=dev))


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


pooled = {}
for name, m in (("ref", ref), ("emu", emu)):
    m.encoder_q[1].register_forward_hook(lambda mod, i, o, name=name: pooled.setdefault(name, []).append(o.detach().flatten(1)))
orig = prod.encoder_q[0].encode
def enc(*a, **k):
    o = orig(*a, **k)
    pooled.setdefault("prod", []).append(o.detach())
    return o
prod.encoder_q[0].encode = enc
outs = {}
for name, m in (("ref", ref), ("emu", emu), ("prod", prod)):
    np.random.seed(11)
    outs[name] = m(x)
for i in range(2):
    print(f"pooled features pass {i}: prod vs ref {rel(pooled['prod'][i], pooled['ref'][i]):.3e}  emu vs ref {rel(pooled['emu'][i], pooled['ref'][i]):.3e}  "
          f"prod vs emu {rel(pooled['prod'][i], pooled['emu'][i]):.3e}")
heads = [n for n, _ in ref.named_parameters() if not n.startswith("encoder_q.0.")]
losses = [k for k in outs["ref"] if "loss" in k]
for k in losses + ["total"]:
    for m in (ref, emu, prod):
        m.zero_grad(set_to_none=True)
    for name, m in (("ref", ref), ("emu", emu), ("prod", prod)):
        l = outs[name][k] if k != "total" else sum(outs[name][q] for q in losses)
        l.backward(retain_graph=True)
    P = dict(prod.named_parameters()); E_ = dict(emu.named_parameters()); R = dict(ref.named_parameters())
    print(f"== {k}: value ref {float(outs['ref'][k]) if k != 'total' else 0:.5f}")
    for n in heads:
        if R[n].grad is None:
            continue
        print(f"   {n:32s} prod/ref {rel(P[n].grad, R[n].grad):.3f}  emu/ref {rel(E_[n].grad, R[n].grad):.3f}  prod/emu {rel(P[n].grad, E_[n].grad):.3f}  |g| {R[n].grad.norm().item():.3e}")
    bb = [n for n in R if n.startswith("encoder_q.0.") and R[n].grad is not None and P[n].grad is not None and E_[n].grad is not None]
    ep = sorted(rel(P[n].grad, R[n].grad) for n in bb); ee = sorted(rel(E_[n].grad, R[n].grad) for n in bb)
    print(f"   backbone median prod/ref {ep[len(ep)//2]:.3f} emu/ref {ee[len(ee)//2]:.3f}")
