"""GPU diagnostic: fp32-mode convolution as ONE launch over all plane products (dv_conv3d_fprop_f32planes /
dv_conv3d_dgrad_f32planes) against one dv_conv3d_*_f32acc launch per product: same sum, time of both."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dualvar_b200 import _lib
from dualvar_b200._lib import ptr, call, stream_ptr
import kernel_handles as K
dev = "cuda:0"
NCL = int(os.environ.get("NCLIPS", "48"))
KP = 3
terms = sorted(((i, j) for i in range(KP) for j in range(KP - i)), key=lambda t: -(t[0] + t[1]))
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (n, t, h, w, ci, co, k, s, p) in [(NCL, 16, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                                      (NCL, 16, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                                      (NCL, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                                      (NCL, 16, 56, 56, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1))]:
    g = K.make_geom(n, t, h, w, ci, co, k, s, p)
    gen = torch.Generator(device=dev).manual_seed(0)
    scale = [1.0, 2.0 ** -8, 2.0 ** -16]
    xp = torch.stack([(torch.randn(n, t, h, w, g.Cin_p, device=dev, generator=gen) * scale[i]).bfloat16() for i in range(KP)])
    xp[..., ci:] = 0
    wfl, wtl = [], []
    for j in range(KP):
        wj = torch.randn(co, ci, *k, device=dev, generator=gen) * scale[j] / 20
        wf, wt = K.pack_conv_weight(wj, g)
        wfl.append(wf); wtl.append(wt)
    wf_all, wt_all = torch.cat(wfl, 1).contiguous(), torch.cat(wtl, 1).contiguous()
    y0 = torch.empty(n, g.To, g.Ho, g.Wo, g.Cout_p, device=dev)
    y1 = torch.full_like(y0, float("nan"))
    def per_product():
        for m, (i, j) in enumerate(terms):
            call("dv_conv3d_fprop_f32acc", ptr(xp[i]), ptr(wfl[j]), ptr(y0), None, ctypes.byref(g), 1 if m else 0, stream_ptr())
    stride1 = True      # strided layers: one tensor map per (stride-parity class, plane)
    merged = lambda: call("dv_conv3d_fprop_f32planes", ptr(xp), xp[0].numel(), KP, ptr(wf_all), ptr(y1), None, None, ctypes.byref(g), stream_ptr())
    per_product()
    line = f"{(ci, co, k, s)}: fprop per-product {timed(per_product):.3f} ms"
    if stride1:
        merged(); torch.cuda.synchronize()
        err = ((y1 - y0).abs().max() / y0.abs().max()).item()
        line += f", merged {timed(merged):.3f} ms (rel diff {err:.1e})"
    dyp = torch.stack([(torch.randn(n, g.To, g.Ho, g.Wo, g.Cout_p, device=dev, generator=gen) * scale[i]).bfloat16() for i in range(KP)])
    dyp[..., co:] = 0
    dx0 = torch.empty(n, t, h, w, g.Cin_p, device=dev); dx1 = torch.full_like(dx0, float("nan"))
    def d_per_product():
        for m, (i, j) in enumerate(terms):
            call("dv_conv3d_dgrad_f32acc", ptr(dyp[i]), ptr(wtl[j]), ptr(dx0), ctypes.byref(g), 1 if m else 0, stream_ptr())
    d_merged = lambda: call("dv_conv3d_dgrad_f32planes", ptr(dyp), dyp[0].numel(), KP, ptr(wt_all), ptr(dx1), ctypes.byref(g), stream_ptr())
    d_per_product(); d_merged(); torch.cuda.synchronize()
    err = ((dx1 - dx0).abs().max() / dx0.abs().max()).item()
    line += f" | dgrad per-product {timed(d_per_product):.3f} ms, merged {timed(d_merged):.3f} ms (rel diff {err:.1e})"
    dw0 = torch.empty(g.Cout_p, g.taps, g.Cin_p, device=dev); dw1 = torch.full_like(dw0, float("nan"))
    def w_per_product():
        for m, (i, j) in enumerate(terms):
            call("dv_conv3d_wgrad_bf16" if m == 0 else "dv_conv3d_wgrad_bf16_acc", ptr(xp[i]), ptr(dyp[j]), ptr(dw0), ctypes.byref(g), stream_ptr())
    w_merged = lambda: call("dv_conv3d_wgrad_f32planes", ptr(xp), ptr(dyp), KP, ptr(dw1), ctypes.byref(g), stream_ptr())
    w_per_product(); w_merged(); torch.cuda.synchronize()
    err = ((dw1 - dw0).abs().max() / dw0.abs().max()).item()
    line += f" | wgrad per-product {timed(w_per_product):.3f} ms, merged {timed(w_merged):.3f} ms (rel diff {err:.1e})"
    print(line, flush=True)
