import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from dualvar_b200 import engine as E, backbones as PB
dev = "cuda:0"
torch.manual_seed(0)
net, _ = PB.select_backbone("r21d"); net = net.to(dev).train()
x = torch.randn(24, 3, 8, 64, 64, device=dev)
log = []
orig_cs, orig_act, orig_ing = E.conv_stats, E.activate, E.ingest
def cs(ctx, xa, conv, bn):
    r = orig_cs(ctx, xa, conv, bn)
    torch.cuda.synchronize()
    log.append(("conv_y", tuple(r.y.shape), r.y.float().abs().sum().item(), r.ss.abs().sum().item() if r.ss is not None else 0))
    return r
def act(ctx, r1, r2=None, res=None, relu=True, out=None, out_coff=0):
    o = orig_act(ctx, r1, r2=r2, res=res, relu=relu, out=out, out_coff=out_coff)
    torch.cuda.synchronize()
    log.append(("act", tuple(o.data.shape), o.data.float().abs().sum().item(), 0))
    return o
def ing(*a, **k):
    o = orig_ing(*a, **k)
    torch.cuda.synchronize()
    log.append(("ingest", tuple(o.data.shape), o.data.float().abs().sum().item(), 0))
    return o
E.conv_stats, E.activate, E.ingest = cs, act, ing
import dualvar_b200.backbones as BB
runs = []
for it in range(4):
    log.clear()
    with torch.no_grad():
        y = net(x)
    runs.append(list(log) + [("out", tuple(y.shape), y.abs().sum().item(), 0)])
for i, rows in enumerate(zip(*runs)):
    base = rows[0]
    bad = [j for j, r in enumerate(rows) if abs(r[2] - base[2]) > 1e-3 * abs(base[2]) + 1e-6 or abs(r[3] - base[3]) > 1e-3 * abs(base[3]) + 1e-6]
    print(i, base[0], base[1], " ".join(f"{r[2]:.6g}/{r[3]:.5g}" for r in rows), "<-- DIFF" if bad else "")
