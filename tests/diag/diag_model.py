"""GPU diagnostic: product modules vs the oracle (fp32 torch on the same GPU), same weights/inputs."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np
import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"

from oracle import models as OM, backbones as OB
from dualvar_b200 import models as PM, backbones as PB


def seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


def relerr(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def cmp_backbone(name, shape=(8, 3, 8, 64, 64)):
    if name in ("s3d", "s3dg"):
        shape = (4, 3, 16, 64, 64)
    seed(0)
    ref, _ = OB.select_backbone(name)
    ref = ref.to(dev).train()
    prod, _ = PB.select_backbone(name)
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev).train()
    x = torch.randn(*shape, device=dev)
    yr = ref(x)
    yp = prod(x)
    print(f"[{name}] fwd rel {relerr(yp, yr):.3e} shape {tuple(yp.shape)}")
    g = torch.randn_like(yr)
    yr.backward(g)
    yp.backward(g)
    worst = ("", 0.0)
    for (n, pr), (_, pp) in zip(ref.named_parameters(), prod.named_parameters()):
        if pp.grad is None:
            print("   no grad for", n); continue
        e = relerr(pp.grad, pr.grad)
        if e > worst[1]:
            worst = (n, e)
    print(f"[{name}] worst param-grad rel err {worst[1]:.3e} at {worst[0]}")
    # yardstick: the oracle itself under bf16 autocast vs its fp32 self
    import copy
    ref2 = copy.deepcopy(ref)
    for p_ in ref2.parameters(): p_.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = ref2(x)
    ya.float().backward(g)
    w2 = ("", 0.0); allp = []; alla = []
    for (n, pr), (_, pa), (_, pp) in zip(ref.named_parameters(), ref2.named_parameters(), prod.named_parameters()):
        e = relerr(pa.grad, pr.grad); alla.append(e); allp.append(relerr(pp.grad, pr.grad))
        if e > w2[1]: w2 = (n, e)
    print(f"[{name}] yardstick autocast-bf16 oracle: fwd rel {relerr(ya, yr):.3e} worst grad rel {w2[1]:.3e} at {w2[0]}; "
          f"median grad rel: autocast {sorted(alla)[len(alla)//2]:.3e} product {sorted(allp)[len(allp)//2]:.3e}")
    # running stats
    wr = 0.0
    for (n, br), (_, bp) in zip(ref.named_buffers(), prod.named_buffers()):
        if br.dtype.is_floating_point:
            wr = max(wr, relerr(bp, br))
        else:
            assert int(br) == int(bp), (n, int(br), int(bp))
    print(f"[{name}] worst running-stat rel err {wr:.3e}")


def cmp_simclr(net, B=8, shape=(8, 64, 64)):
    args = SimpleNamespace(shufflerank_theta=0.05)
    seed(0)
    ref = OM.SimCLR_TimeSeriesV4(net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args).to(dev).train()
    prod = PM.SimCLR_TimeSeriesV4(net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev).train()
    x = torch.randn(B, 3, 3, *shape, device=dev)
    np.random.seed(5); rr = ref(x)
    np.random.seed(5); rp = prod(x)
    for k in rr:
        if "labels" in k: continue
        print(f"[simclr {net}] {k:40s} ref {rr[k].float().mean().item():+.5f} prod {rp[k].float().mean().item():+.5f} rel {relerr(rp[k], rr[k]):.3e}")
    lr = sum(v for k, v in rr.items() if "loss" in k)
    lp = sum(v for k, v in rp.items() if "loss" in k)
    lr.backward(); lp.backward()
    print(f"[simclr {net}] total ref {lr.item():.6f} prod {lp.item():.6f}")
    errs = []
    for (n, pr), (_, pp) in zip(ref.named_parameters(), prod.named_parameters()):
        if pp.grad is None:
            print("   no grad for", n); continue
        errs.append((relerr(pp.grad, pr.grad), n, pr.grad.norm().item()))
    errs.sort(reverse=True)
    for e, n, gn in errs[:6]:
        print(f"   grad rel {e:.3e} |g|={gn:.3e} {n}")
    print(f"   median grad rel {errs[len(errs)//2][0]:.3e}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["r21d", "r3d", "simclr"]
    for w in which:
        try:
            if w in ("r21d", "r3d", "c3d", "s3d", "s3dg"):
                cmp_backbone(w)
            elif w == "simclr":
                cmp_simclr("r21d"); cmp_simclr("r3d")
        except Exception as ex:
            import traceback; traceback.print_exc()
            try: torch.cuda.synchronize()
            except Exception as e2: print("device dead", e2); break
