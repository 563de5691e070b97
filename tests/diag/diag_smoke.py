import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from dualvar_b200 import models as PM
from oracle import models as OM
dev = "cuda:0"
args = SimpleNamespace(shufflerank_theta=0.05)
torch.manual_seed(0); np.random.seed(0); random.seed(0)
ref = OM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args).to(dev).train()
prod = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", args)
prod.load_state_dict(ref.state_dict()); prod = prod.to(dev).train()
x = torch.randn(8, 3, 3, 8, 64, 64, device=dev)
for seed in (1, 5, 1, 2):
    np.random.seed(seed); perms = np.array([np.random.permutation(2) for _ in range(8)])
    np.random.seed(seed); rr = ref(x)
    np.random.seed(seed); rp = prod(x)
    np.random.seed(seed); rp2 = prod(x)
    print("seed", seed, "perms", perms.tolist())
    for k in rr:
        if "loss" in k:
            print(f"   {k:40s} ref {rr[k].item():.5f} prod {rp[k].item():.5f} prod-again {rp2[k].item():.5f}")
