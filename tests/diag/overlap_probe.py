"""GPU diagnostic: do a conv_tile_kernel launch (one stream) and a BatchNorm pass (another stream) overlap on B200?"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests helpers (kernel_handles)
import torch
from dualvar_b200 import _lib
import kernel_handles as K
from dualvar_b200._lib import ptr, call
dev = "cuda:0"
n, t, h, w, ci, co = 96, 16, 56, 56, 64, 144
g = K.make_geom(n, t, h, w, ci, co, (1, 3, 3), (1, 1, 1), (0, 1, 1))
x = torch.randn(n, t, h, w, g.Cin_p, device=dev).bfloat16()
wt = torch.randn(co, ci, 1, 3, 3, device=dev) / 20
wf, wtt = K.pack_conv_weight(wt, g)
y = torch.empty(n, t, h, w, g.Cout_p, device=dev, dtype=torch.bfloat16)
dy = torch.randn_like(y)
dwp = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.float32, device=dev)
stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
rows, Cp = n * t * h * w, 144
yb = torch.randn(rows, Cp, device=dev).bfloat16(); zb = torch.empty_like(yb); db = torch.randn_like(yb); dyb = torch.empty_like(yb)
ss = torch.randn(2 * Cp, device=dev); coef = torch.randn(3 * Cp, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
sp = lambda s: ctypes.c_void_p(s.cuda_stream)
convs = {"fprop": lambda s: call("dv_conv3d_fprop_bf16", ptr(x), ptr(wf), ptr(y), ptr(stats), None, ctypes.byref(g), sp(s)),
         "dgrad": lambda s: call("dv_conv3d_dgrad_bf16", ptr(dy), ptr(wtt), ptr(x), ctypes.byref(g), sp(s)),
         "wgrad": lambda s: call("dv_conv3d_wgrad_bf16", ptr(x), ptr(dy), ptr(dwp), ctypes.byref(g), sp(s))}
bns = {"bn_apply": lambda s: call("dv_bn_apply", ptr(yb), ptr(ss), None, None, None, ptr(zb), rows, Cp, Cp, 0, 1, sp(s)),
       "bn_bwd_apply": lambda s: call("dv_bn_bwd_apply", ptr(db), None, ptr(zb), ptr(yb), ptr(ss), ptr(coef), ptr(dyb), None, rows, Cp, Cp, 0, 1, sp(s)),
       "bn_bwd_reduce": lambda s: call("dv_bn_bwd_reduce", ptr(db), None, ptr(zb), ptr(yb), ptr(ss), ptr(stats), rows, Cp, Cp, 0, 1, sp(s))}
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cur = torch.cuda.current_stream()
for cn, cf in convs.items():
    for bn, bf in bns.items():
        def seq():
            cf(cur); bf(cur)
        def par():
            s1.wait_stream(cur); s2.wait_stream(cur)
            cf(s1); bf(s2)
            cur.wait_stream(s1); cur.wait_stream(s2)
        def par_rev():
            s1.wait_stream(cur); s2.wait_stream(cur)
            bf(s2); cf(s1)
            cur.wait_stream(s1); cur.wait_stream(s2)
        tc, tb = timeit(lambda: cf(cur)), timeit(lambda: bf(cur))
        print(f"{cn} {tc:.3f} ms + {bn} {tb:.3f} ms: sequential {timeit(seq):.3f}, two streams (conv first) {timeit(par):.3f}, (bn first) {timeit(par_rev):.3f}", flush=True)
