"""Target for `ncu --set full --import-source on`: one launch each of the weakest conv launches of the bench step at
bench size (192 clips): stem fprop / wgrad (space-to-depth 7x7 stride 2), stride-2 dgrad / wgrad 64->230 3x3."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dualvar_b200 import _lib
import kernel_handles as K
from dualvar_b200._lib import ptr, call, stream_ptr
dev = "cuda:0"
n = int(os.environ.get("NCLIPS", "192"))
# stem
gs = K.make_geom(n, 16, 112, 112, 3, 83, (1, 7, 7), (1, 2, 2), (0, 3, 3))
xs = torch.randn(n, 16, 56, 59, 16, device=dev).bfloat16()
ws = (torch.randn(gs.Cout_p, 4, 64, device=dev) / 20).bfloat16()
ys = torch.empty(n, 16, 56, 56, gs.Cout_p, device=dev, dtype=torch.bfloat16)
dys = torch.randn_like(ys)
stats = torch.zeros(2 * gs.Cout_p, dtype=torch.float64, device=dev)
dws = torch.empty((gs.Cout_p, 4, 64), dtype=torch.float32, device=dev)
# stride-2 spatial conv 64 -> 230
g2 = K.make_geom(n, 16, 56, 56, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1))
x2 = torch.randn(n, 16, 56, 56, 64, device=dev).bfloat16()
w2 = torch.randn(230, 64, 1, 3, 3, device=dev) / 20
wf2, wt2 = K.pack_conv_weight(w2, g2)
dy2 = torch.randn(n, 16, 28, 28, g2.Cout_p, device=dev).bfloat16()
dx2 = torch.empty_like(x2)
dw2 = torch.empty((g2.Cout_p, 9, 64), dtype=torch.float32, device=dev)


def run():
    call("dv_conv3d_stem_fprop_bf16", ptr(xs), ptr(ws), ptr(ys), ptr(stats), None, ctypes.byref(gs), stream_ptr())
    call("dv_conv3d_stem_wgrad_bf16", ptr(xs), ptr(dys), ptr(dws), ctypes.byref(gs), stream_ptr())
    call("dv_conv3d_dgrad_bf16", ptr(dy2), ptr(wt2), ptr(dx2), ctypes.byref(g2), stream_ptr())
    call("dv_conv3d_wgrad_bf16", ptr(x2), ptr(dy2), ptr(dw2), ctypes.byref(g2), stream_ptr())
    torch.cuda.synchronize()


run(); run()
print("ok")
