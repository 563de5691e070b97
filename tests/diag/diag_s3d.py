import os, sys, random, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import backbones as OB
from dualvar_b200 import backbones as PB
dev = "cuda:0"
def rel(a, b): return ((a.float()-b.float()).norm()/(b.float().norm()+1e-12)).item()
for name in ("s3d", "s3dg"):
    torch.manual_seed(0)
    ref, _ = OB.select_backbone(name); ref = ref.to(dev)
    prod, _ = PB.select_backbone(name); prod.load_state_dict(ref.state_dict()); prod = prod.to(dev)
    # make BN non-trivial but well conditioned for eval: random running stats
    x = torch.randn(4, 3, 16, 64, 64, device=dev)
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.momentum = 1.0
                m.weight.uniform_(0.8, 1.2); m.bias.normal_(0, 0.1)
        ref.train(); ref(x)          # running stats := batch stats of this input
    prod.load_state_dict(ref.state_dict())
    ref.eval(); prod.eval()
    with torch.no_grad():
        yr = ref(x); yp = prod(x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ya = ref(x)
    print(f"[{name}] eval fwd rel: product {rel(yp, yr):.3e} autocast-yardstick {rel(ya, yr):.3e} shape {tuple(yp.shape)} |y| {yr.abs().mean().item():.3e}")
    # train mode, bigger batch/extent
    ref.train(); prod.train()
    x = torch.randn(8, 3, 16, 128, 128, device=dev)
    yr = ref(x); yp = prod(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = copy.deepcopy(ref)(x)
    print(f"[{name}] train fwd rel: product {rel(yp, yr):.3e} autocast-yardstick {rel(ya, yr):.3e} shape {tuple(yp.shape)}")
