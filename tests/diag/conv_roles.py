"""GPU diagnostic: where do the producer / MMA / epilogue roles of conv_tile_kernel spend their cycles?"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
# the dv_debug_* entry points live in the diagnostics build only (`make diag`)
os.environ.setdefault("DV_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dualvar_b200", "lib", "libdualvar_b200_diag.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests helpers (kernel_handles)
import torch
from dualvar_b200 import _lib
import kernel_handles as K
dev = "cuda:0"
N = int(os.environ.get("NCLIPS", "96"))
LAYERS = [("fprop temporal 144->64", "f", (N, 16, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))),
          ("fprop spatial 64->144", "f", (N, 16, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("dgrad spatial 64->144", "d", (N, 16, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("dgrad temporal 144->64", "d", (N, 16, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))),
          ("dgrad+bnred temporal 144->64", "r", (N, 16, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))),
          ("dgrad+bnred spatial 64->144", "r", (N, 16, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("fprop temporal 83->64", "f", (N, 16, 56, 56, 83, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))),
          ("fprop spatial 128->288", "f", (N, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("dgrad spatial s2 64->230", "d", (N, 16, 56, 56, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1))),
          ("stem fprop 3->83 7x7 s2", "s", (N, 16, 112, 112, 3, 83, (1, 7, 7), (1, 2, 2), (0, 3, 3)))]
if os.environ.get("ONLY"):
    LAYERS = [l for l in LAYERS if any(k in l[0] for k in os.environ["ONLY"].split(","))]
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
for name, kind, (n, t, h, w, ci, co, k, s, p) in LAYERS:
    g = K.make_geom(n, t, h, w, ci, co, k, s, p)
    x = torch.randn(n, t, h, w, g.Cin_p, device=dev).bfloat16()
    wt = torch.randn(co, ci, *k, device=dev) / 20
    wf, wtt = K.pack_conv_weight(wt, g)
    dy = torch.randn(n, g.To, g.Ho, g.Wo, g.Cout_p, device=dev).bfloat16()
    stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
    ss = torch.randn(2 * g.Cin_p, device=dev)
    sums = torch.zeros(2 * g.Cin_p, dtype=torch.float64, device=dev)
    dxb = torch.empty_like(x)
    if kind == "s":
        xs = torch.randn(n, t, h // 2, w // 2 + 3, 16, device=dev).bfloat16()
        ws = (torch.randn(g.Cout_p, k[0] * 4, 64, device=dev) / 20).bfloat16()
        ys = torch.empty(n, g.To, g.Ho, g.Wo, g.Cout_p, device=dev, dtype=torch.bfloat16)
    run = {"f": lambda: K.conv3d_fprop(x, wf, g, bn_stats=stats), "d": lambda: K.conv3d_dgrad(dy, wtt, g),
           "s": lambda: _lib.call("dv_conv3d_stem_fprop_bf16", _lib.ptr(xs), _lib.ptr(ws), _lib.ptr(ys), _lib.ptr(stats), None,
                                  ctypes.byref(g), _lib.stream_ptr()),
           "r": lambda: _lib.call("dv_conv3d_dgrad_bnred_bf16", _lib.ptr(dy), _lib.ptr(wtt), _lib.ptr(dxb), ctypes.byref(g),
                                  _lib.ptr(x), _lib.ptr(ss), _lib.ptr(sums), _lib.stream_ptr())}[kind]
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof.zero_()
    _lib.load().dv_debug_set_conv_profile(ctypes.c_void_p(prof.data_ptr()))
    run(); torch.cuda.synchronize()
    _lib.load().dv_debug_set_conv_profile(None)
    P = prof.view(148, 16).double()
    tiles = P[:, 7].clamp_min(1)
    live = P[:, 7] > 0
    per = lambda i: (P[live, i] / tiles[live]).mean().item()
    lead = P[:, 2] > 0          # CTAs that issued MMAs (every CTA, or the pair leaders)
    perl = lambda i: (P[lead, i] / tiles[lead]).mean().item()
    fl = 2.0 * n * g.To * g.Ho * g.Wo * co * ci * k[0] * k[1] * k[2]
    print(f"{name}: {ms:.3f} ms {fl/ms/1e9:.0f} TF/s | cycles/tile: producer {per(0):.0f} (wait-free-stage {per(1):.0f}) | "
          f"mma {perl(2):.0f} (wait-data {perl(3):.0f}, wait-acc {perl(4):.0f}) | epilogue {per(5):.0f} (wait-acc {per(6):.0f}: y-ld {per(8):.0f} "
          f"tmem-ld {per(9):.0f} cvt+sts {per(10):.0f} bar {per(11):.0f} store {per(12):.0f} colsum {per(13):.0f}) | tiles/CTA {tiles.mean().item():.0f}", flush=True)
