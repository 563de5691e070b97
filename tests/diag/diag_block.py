import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests helpers (kernel_handles)
import torch, torch.nn as nn, torch.nn.functional as F
from dualvar_b200 import engine as E
import kernel_handles as K
torch.backends.cudnn.allow_tf32 = False
dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(4)
N, C, T, H, W = 6, 64, 4, 16, 16
conv = nn.Conv3d(C, C, (1, 3, 3), padding=(0, 1, 1), bias=False).to(dev)
bn = nn.BatchNorm3d(C).to(dev)
bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.2)
import copy
conv_r, bn_r = copy.deepcopy(conv), copy.deepcopy(bn)
conv_r.weight.data = conv_r.weight.data.bfloat16().float()
x = torch.randn(N, C, T, H, W, device=dev, generator=gen).bfloat16().float()
for use_res in (False, True):
    xr = x.clone().requires_grad_(True)
    pre = bn_r(conv_r(xr))
    yr = F.relu(pre + xr) if use_res else F.relu(pre)
    gy = torch.randn(yr.shape, device=dev, generator=gen).bfloat16().float()
    yr.backward(gy)
    ctx = E.Context(training=True)
    xa = E.Act(K.to_ndhwc(x), C)
    out = E.activate(ctx, E.conv_stats(ctx, xa, conv, bn), res=xa if use_res else None)
    y = K.from_ndhwc(out.data, C)
    out.grad = K.to_ndhwc(gy)
    E.run_backward(ctx)
    dx = K.from_ndhwc(xa.grad, C)
    err = (dx - xr.grad).abs()
    print("res", use_res, "fwd maxerr", (y - yr).abs().max().item(), "dx maxerr", err.max().item(), "max|g|", xr.grad.abs().max().item(),
          "frac>0.05", (err > 0.05).float().mean().item(), "argmax", torch.nonzero(err == err.max())[0].tolist())
    # mask disagreement: positions where relu mask differs (y near 0)
    m_ref = (yr > 0); m_our = (y > 0)
    print("   mask disagreements", (m_ref != m_our).float().mean().item())
    bad = err > 0.05
    print("   of bad elems, fraction where masks disagree:", ((m_ref != m_our) & bad).float().sum().item() / max(1, bad.float().sum().item()))
    conv_r.weight.grad = None; bn_r.weight.grad = None; bn_r.bias.grad = None
