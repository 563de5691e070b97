"""fp32 mode at the bench geometry (R(2+1)D-18 SimCLR+DualVar, 16x112x112): step time next to the bf16 mode, the
torch fp32 (cuDNN, TF32 off) oracle step on the same GPU, and the loss agreement with that oracle."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
import numpy as np, torch
from dualvar_b200 import engine as E, models as PM
from dualvar_b200.optim import SGD
from oracle import models as OM
dev = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ARGS = SimpleNamespace(shufflerank_theta=0.05)
B = int(os.environ.get("B", "16"))
a = ("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS)
torch.manual_seed(0); np.random.seed(0); random.seed(0)
ref = OM.SimCLR_TimeSeriesV4(*a).to(dev).train()
state = {k: v.clone() for k, v in ref.state_dict().items()}
x = torch.randn(B, 3, 3, 16, 112, 112, device=dev)


def timed(model, opt, tag):
    def step():
        np.random.seed(7)
        ret = model(x); loss = sum(v for k, v in ret.items() if "loss" in k)
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); return ret
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{tag}: B={B} 16x112x112  {ms:.1f} ms/step  {B / ms * 1e3:.1f} samples/s", flush=True)


# losses of the first step from identical weights
np.random.seed(7); rr = ref(x)
for mode, planes in (("fp32", 3), ("fp32", 2), ("bf16", None)):
    E.set_precision(mode, planes)
    m = PM.SimCLR_TimeSeriesV4(*a)
    m.load_state_dict(state)
    m = m.to(dev).train()
    np.random.seed(7); rp = m(x)
    errs = {k: abs(rp[k].item() - rr[k].item()) / abs(rr[k].item()) for k in rr if "loss" in k}
    print(f"{mode}{'' if planes is None else f' ({planes} planes)'} loss rel err vs torch fp32: " +
          ", ".join(f"{k.replace('_contrast_loss', '')} {v:.2e}" for k, v in errs.items()), flush=True)
    del rp
    timed(m, SGD([{'params': p} for p in m.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9),
          f"dualvar_b200 {mode}{'' if planes is None else f' ({planes} planes)'}")
    del m; torch.cuda.empty_cache()
E.set_precision("bf16")
del rr
timed(ref, torch.optim.SGD(ref.parameters(), lr=0.003, weight_decay=1e-4, momentum=0.9), "torch fp32 oracle (cuDNN, TF32 off)")


# the "honest bar" of SURVEY §8(d): the same torch modules on this B200 through cuDNN under bf16 autocast
class _Autocast(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, x):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return {k: (v.float() if v.is_floating_point() else v) for k, v in self.m(x).items()}


ref.load_state_dict(state)
timed(_Autocast(ref), torch.optim.SGD(ref.parameters(), lr=0.003, weight_decay=1e-4, momentum=0.9),
      "torch bf16 autocast oracle (cuDNN)")
torch.backends.cudnn.benchmark = True
timed(_Autocast(ref), torch.optim.SGD(ref.parameters(), lr=0.003, weight_decay=1e-4, momentum=0.9),
      "torch bf16 autocast oracle (cuDNN, cudnn.benchmark)")
refcl = ref.to(memory_format=torch.channels_last_3d)
timed(_Autocast(refcl), torch.optim.SGD(refcl.parameters(), lr=0.003, weight_decay=1e-4, momentum=0.9),
      "torch bf16 autocast oracle (cuDNN, channels_last_3d, cudnn.benchmark)")
