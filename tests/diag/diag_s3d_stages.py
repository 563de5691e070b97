import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import backbones as OB
from dualvar_b200 import backbones as PB, engine as E, s3dg as PS
dev = "cuda:0"
def rel(a, b): return ((a.float()-b.float()).norm()/(b.float().norm()+1e-12)).item()
for name in ("s3dg",):
    torch.manual_seed(0)
    ref, _ = OB.select_backbone(name); ref = ref.to(dev).train()
    prod, _ = PB.select_backbone(name); prod.load_state_dict(ref.state_dict()); prod = prod.to(dev).train()
    x = torch.randn(8, 3, 16, 64, 64, device=dev)
    names = ["Conv_1a", "MaxPool_2a", "Conv_2b", "Conv_2c", "MaxPool_3a", "Mixed_3b", "Mixed_3c", "MaxPool_4a", "Mixed_4b", "Mixed_4c",
             "Mixed_4d", "Mixed_4e", "Mixed_4f", "MaxPool_5a", "Mixed_5b", "Mixed_5c"]
    ref_out = {}
    hooks = []
    seen = set()
    for n in names:
        m = getattr(ref, n)
        def mk(n):
            def h(mod, inp, out):
                if n not in ref_out: ref_out[n] = out.detach()
            return h
        hooks.append(m.register_forward_hook(mk(n)))
    with torch.no_grad():
        ref(x)
        ref_a = {}
        # autocast yardstick per stage
        hs = []
        for n in names:
            def mk2(n):
                def h(mod, inp, out):
                    if n not in ref_a: ref_a[n] = out.detach().float()
                return h
            hs.append(getattr(ref, n).register_forward_hook(mk2(n)))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ref(x)
    prod_out = {}
    for n in names:
        m = getattr(prod, n)
        if hasattr(m, "run"):
            orig = m.run
            def mk3(n, orig):
                def run(ctx, xa, *a, **k):
                    r = orig(ctx, xa, *a, **k)
                    act = r[0] if isinstance(r, tuple) else r
                    prod_out[n] = E.to_ncdhw(act)
                    return r
                return run
            m.run = mk3(n, orig)
    orig_pool = PS.S3D._pool
    def pool(ctx, xa, mod):
        o = orig_pool(ctx, xa, mod)
        for n in names:
            if getattr(prod, n) is mod: prod_out[n] = E.to_ncdhw(o)
        return o
    PS.S3D._pool = staticmethod(pool)
    with torch.no_grad():
        prod(x)
    for n in names:
        if n in prod_out:
            print(f"{n:12s} product {rel(prod_out[n], ref_out[n]):.3e}  autocast {rel(ref_a[n], ref_out[n]):.3e}  shape {tuple(ref_out[n].shape)}")
