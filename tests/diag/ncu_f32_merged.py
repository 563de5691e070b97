"""Target for `ncu --set full`: the fp32 mode's merged plane-product launches of the layer-1 3x3 convolution 64->144 at
48 clips (B = 16), 3 planes: fprop (batch statistics in the epilogue), dgrad (two-issuer instance) and wgrad. One warm-up
of each, then one profiled launch of each (-k regex:conv_ --launch-skip 3 -c 3)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dualvar_b200._lib import ptr, call, stream_ptr
import kernel_handles as K
dev = "cuda:0"
KP = 3
n, t, h, w, ci, co = 48, 16, 56, 56, 64, 144
g = K.make_geom(n, t, h, w, ci, co, (1, 3, 3), (1, 1, 1), (0, 1, 1))
gen = torch.Generator(device=dev).manual_seed(0)
scale = [1.0, 2.0 ** -8, 2.0 ** -16]
xp = torch.stack([(torch.randn(n, t, h, w, g.Cin_p, device=dev, generator=gen) * scale[i]).bfloat16() for i in range(KP)])
packs = [K.pack_conv_weight(torch.randn(co, ci, 1, 3, 3, device=dev, generator=gen) * scale[j] / 20, g) for j in range(KP)]
wf_all = torch.cat([p[0] for p in packs], 1).contiguous()
wt_all = torch.cat([p[1] for p in packs], 1).contiguous()
y = torch.empty(n, t, h, w, g.Cout_p, device=dev)
stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
dyp = torch.stack([(torch.randn(n, t, h, w, g.Cout_p, device=dev, generator=gen) * scale[i]).bfloat16() for i in range(KP)])
dx = torch.empty(n, t, h, w, g.Cin_p, device=dev)
dw = torch.empty(g.Cout_p, g.taps, g.Cin_p, device=dev)
def run():
    call("dv_conv3d_fprop_f32planes", ptr(xp), xp.stride(0), KP, ptr(wf_all), ptr(y), ptr(stats), None, ctypes.byref(g), stream_ptr())
    call("dv_conv3d_dgrad_f32planes", ptr(dyp), dyp.stride(0), KP, ptr(wt_all), ptr(dx), ctypes.byref(g), stream_ptr())
    call("dv_conv3d_wgrad_f32planes", ptr(xp), ptr(dyp), KP, ptr(dw), ctypes.byref(g), stream_ptr())
    torch.cuda.synchronize()
run(); run()
print("ok")
