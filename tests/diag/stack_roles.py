"""GPU diagnostic (diagnostics build): role counters of the kh-stacked dgrad."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("DV_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dualvar_b200", "lib", "libdualvar_b200_diag.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dualvar_b200 import _lib
import kernel_handles as K
dev = "cuda:0"
n, t, h, w, ci, co = 192, 16, 56, 56, 64, int(os.environ.get("CO", "144"))
g = K.make_geom(n, t, h, w, ci, co, (1, 3, 3), (1, 1, 1), (0, 1, 1))
wt_f = torch.randn(co, ci, 1, 3, 3, device=dev) / 20
_, wt = K.pack_conv_weight(wt_f, g)
ws = wt.view(g.Cin_p, 3, 3, g.Cout_p).flip(1).permute(1, 0, 2, 3).reshape(3 * g.Cin_p, 3, g.Cout_p).contiguous()
dy = torch.randn(n, t, h, w, g.Cout_p, device=dev).bfloat16()
dx = torch.empty(n, t, h, w, g.Cin_p, device=dev, dtype=torch.bfloat16)
run = lambda: _lib.call("dv_conv3d_dgrad_stack_bf16", _lib.ptr(dy), _lib.ptr(ws), _lib.ptr(dx), ctypes.byref(g), None, None, None, _lib.stream_ptr())
run(); run(); torch.cuda.synchronize()
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
_lib.load().dv_debug_set_conv_profile(ctypes.c_void_p(prof.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
_lib.load().dv_debug_set_conv_profile(None)
P = prof.view(148, 16).double()
lead = P[:, 2] > 0
tiles = 192 * 16 * 4 * 7 / 2 / 74
print(f"{e0.elapsed_time(e1):.3f} ms; per tile: producer {P[:,0].mean().item()/tiles:.0f} (wait-free-stage {P[:,1].mean().item()/tiles:.0f}) | "
      f"mma {P[lead,2].mean().item()/tiles:.0f} (wait-data {P[lead,3].mean().item()/tiles:.0f}, wait-acc {P[lead,4].mean().item()/tiles:.0f})")
