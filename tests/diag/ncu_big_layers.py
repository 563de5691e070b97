"""Target for `ncu --set full`: one launch each of the layer-1 kernels at the bench size (192 clips, 16x56x56):
conv fprop / dgrad / wgrad 64->144 3x3 and the three BatchNorm passes on 9.63 M x 144. First a warm-up of each."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests helpers (kernel_handles)
import torch
from dualvar_b200 import _lib
import kernel_handles as K
from dualvar_b200._lib import ptr, call, stream_ptr
dev = "cuda:0"
n, t, h, w, ci, co = 192, 16, 56, 56, 64, 144
g = K.make_geom(n, t, h, w, ci, co, (1, 3, 3), (1, 1, 1), (0, 1, 1))
x = torch.randn(n, t, h, w, g.Cin_p, device=dev).bfloat16()
wt = torch.randn(co, ci, 1, 3, 3, device=dev) / 20
wf, wtt = K.pack_conv_weight(wt, g)
y = torch.empty(n, t, h, w, g.Cout_p, device=dev, dtype=torch.bfloat16)
dy = torch.randn_like(y); dx = torch.empty_like(x)
dwp = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.float32, device=dev)
stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
rows, Cp = n * t * h * w, 144
z = torch.empty_like(y); dyb = torch.empty_like(y)
ss = torch.randn(2 * Cp, device=dev); coef = torch.randn(3 * Cp, device=dev)
def run():
    call("dv_conv3d_fprop_bf16", ptr(x), ptr(wf), ptr(y), ptr(stats), None, ctypes.byref(g), stream_ptr())
    call("dv_bn_apply", ptr(y), ptr(ss), None, None, None, ptr(z), rows, Cp, Cp, 0, 1, stream_ptr())
    call("dv_bn_bwd_reduce", ptr(dy), None, ptr(z), ptr(y), ptr(ss), ptr(stats), rows, Cp, Cp, 0, 1, stream_ptr())
    call("dv_bn_bwd_apply", ptr(dy), None, ptr(z), ptr(y), ptr(ss), ptr(coef), ptr(dyb), None, rows, Cp, Cp, 0, 1, stream_ptr())
    call("dv_conv3d_wgrad_bf16", ptr(x), ptr(dyb), ptr(dwp), ctypes.byref(g), stream_ptr())
    call("dv_conv3d_dgrad_bf16", ptr(dyb), ptr(wtt), ptr(dx), ctypes.byref(g), stream_ptr())
    torch.cuda.synchronize()
run(); run()
print("ok")
