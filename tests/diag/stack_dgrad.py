"""GPU diagnostic: the kh-stacked data gradient (dv_conv3d_dgrad_stack_bf16) against dv_conv3d_dgrad_bf16 - results and time."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dualvar_b200 import _lib
import kernel_handles as K
dev = "cuda:0"
N = int(os.environ.get("NCLIPS", "192"))
for (n, t, h, w, ci, co) in [(N, 16, 56, 56, 64, 144), (N, 16, 56, 56, 64, 128), (N, 16, 56, 56, 64, 192)]:
    g = K.make_geom(n, t, h, w, ci, co, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    if not _lib.load().dv_conv3d_dgrad_stack_ok(ctypes.byref(g)):
        print("not eligible", (n, t, h, w, ci, co)); continue
    gen = torch.Generator(device=dev).manual_seed(1)
    wt_f = torch.randn(co, ci, 1, 3, 3, device=dev, generator=gen) / 20
    _, wt = K.pack_conv_weight(wt_f, g)
    ws = wt.view(g.Cin_p, 3, 3, g.Cout_p).flip(1).permute(1, 0, 2, 3).reshape(3 * g.Cin_p, 3, g.Cout_p).contiguous()
    dy = torch.randn(n, t, h, w, g.Cout_p, device=dev, generator=gen).bfloat16()
    dy[..., co:] = 0
    ref = K.conv3d_dgrad(dy, wt, g)
    dx = torch.full_like(ref, float("nan"))
    _lib.call("dv_conv3d_dgrad_stack_bf16", _lib.ptr(dy), _lib.ptr(ws), _lib.ptr(dx), ctypes.byref(g), None, None, None, _lib.stream_ptr())
    torch.cuda.synchronize()
    diff = (dx.float() - ref.float()).abs()
    print((n, t, h, w, ci, co), "max |diff|", diff.max().item(), "max |ref|", ref.float().abs().max().item(),
          "nan", int(torch.isnan(dx.float()).sum()), "frac differing", (diff > 0).float().mean().item(), flush=True)
    # fused BatchNorm-backward reduce
    y_prev = torch.randn(n, t, h, w, g.Cin_p, device=dev, generator=gen).bfloat16()
    ss = torch.randn(2 * g.Cin_p, device=dev, generator=gen)
    s_ref = torch.zeros(2 * g.Cin_p, dtype=torch.float64, device=dev)
    s_new = torch.zeros_like(s_ref)
    dx1 = torch.empty_like(ref); dx2 = torch.empty_like(ref)
    _lib.call("dv_conv3d_dgrad_bnred_bf16", _lib.ptr(dy), _lib.ptr(wt), _lib.ptr(dx1), ctypes.byref(g), _lib.ptr(y_prev), _lib.ptr(ss), _lib.ptr(s_ref), _lib.stream_ptr())
    _lib.call("dv_conv3d_dgrad_stack_bf16", _lib.ptr(dy), _lib.ptr(ws), _lib.ptr(dx2), ctypes.byref(g), _lib.ptr(y_prev), _lib.ptr(ss), _lib.ptr(s_new), _lib.stream_ptr())
    torch.cuda.synchronize()
    print("   bnred sums rel diff", ((s_new - s_ref).abs().max() / s_ref.abs().max()).item(), "dx diff", (dx2.float() - dx1.float()).abs().max().item(), flush=True)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    fa = lambda: K.conv3d_dgrad(dy, wt, g)
    fb = lambda: _lib.call("dv_conv3d_dgrad_stack_bf16", _lib.ptr(dy), _lib.ptr(ws), _lib.ptr(dx), ctypes.byref(g), None, None, None, _lib.stream_ptr())
    fc = lambda: _lib.call("dv_conv3d_dgrad_bnred_bf16", _lib.ptr(dy), _lib.ptr(wt), _lib.ptr(dx1), ctypes.byref(g), _lib.ptr(y_prev), _lib.ptr(ss), _lib.ptr(s_ref), _lib.stream_ptr())
    fd = lambda: _lib.call("dv_conv3d_dgrad_stack_bf16", _lib.ptr(dy), _lib.ptr(ws), _lib.ptr(dx2), ctypes.byref(g), _lib.ptr(y_prev), _lib.ptr(ss), _lib.ptr(s_new), _lib.stream_ptr())
    for f in (fa, fb, fc, fd): f(); f()
    ts = [[], [], [], []]
    for _ in range(7):
        for i, f in enumerate((fa, fb, fc, fd)):
            ts[i].append(timed(f))
    med = [sorted(x)[3] for x in ts]
    fl = 2.0 * n * t * h * w * co * ci * 9 / 1e9
    print(f"   dgrad {med[0]:.3f} ms {fl/med[0]:.0f} TF/s | stacked {med[1]:.3f} ms {fl/med[1]:.0f} TF/s | dgrad+bnred {med[2]:.3f} | stacked+bnred {med[3]:.3f}", flush=True)
