"""GPU diagnostic: interleaved A/B timing of the bench step's big conv launches across library builds / environment
toggles in ONE process (clock drift between runs on a box is larger than most kernel changes).
  AB="label=lib.so[,ENV=VAL...];label2=..."   lib names relative to dualvar_b200/lib; default: product library only
Every configuration gets its own dlopen of a private copy of its library (own statics, own read of the environment)."""
import os, sys, ctypes, shutil, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dualvar_b200 import _lib
import kernel_handles as K
dev = "cuda:0"
N = int(os.environ.get("NCLIPS", "192"))
REPS = int(os.environ.get("REPS", "7"))
LIBDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dualvar_b200", "lib")
LAYERS = [("spatial 64->144", (N, 16, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("temporal 144->64", (N, 16, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))),
          ("temporal 83->64", (N, 16, 56, 56, 83, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))),
          ("spatial s2 64->230", (N, 16, 56, 56, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1))),
          ("temporal s2 230->128", (N, 16, 28, 28, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0))),
          ("spatial 128->288", (N, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("temporal 288->128", (N, 8, 28, 28, 288, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0))),
          ("spatial s2 128->460", (N, 8, 28, 28, 128, 460, (1, 3, 3), (1, 2, 2), (0, 1, 1))),
          ("spatial 256->576", (N, 4, 14, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("spatial 512->1152", (N, 2, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("stem", (N, 16, 112, 112, 3, 83, (1, 7, 7), (1, 2, 2), (0, 3, 3))),
          # probes (not layers of the network): the same spatial conv without a partial last K chunk / with fewer channels
          ("probe spatial 64->128", (N, 16, 56, 56, 64, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("probe spatial 64->192", (N, 16, 56, 56, 64, 192, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
          ("probe temporal 128->64", (N, 16, 56, 56, 128, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)))]
if os.environ.get("ONLY"):
    LAYERS = [l for l in LAYERS if any(k in l[0] for k in os.environ["ONLY"].split(","))]

configs = []
tmpdir = tempfile.mkdtemp()
for i, spec in enumerate(os.environ.get("AB", "product=libdualvar_b200.so").split(";")):
    label, rest = spec.split("=", 1)
    parts = rest.split(",")
    env = dict(p.split("=", 1) for p in parts[1:])
    private = os.path.join(tmpdir, f"cfg{i}.so")
    shutil.copy(os.path.join(LIBDIR, parts[0]), private)
    lib = ctypes.CDLL(private)
    for name, (res, args) in _lib._SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    configs.append((label, lib, env))


def use(cfg):
    _lib._lib = cfg[1]
    _lib._fn_cache.clear()


# first conv call of every instance reads its environment
g0 = K.make_geom(2, 4, 16, 16, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1))
x0 = torch.randn(2, 4, 16, 16, 64, device=dev).bfloat16()
w0 = K.pack_conv_weight(torch.randn(64, 64, 1, 3, 3, device=dev), g0)
for cfg in configs:
    saved = {k: os.environ.get(k) for k in cfg[2]}
    os.environ.update(cfg[2])
    use(cfg)
    K.conv3d_fprop(x0, w0[0], g0)
    torch.cuda.synchronize()
    for k, v in saved.items():
        if v is None: os.environ.pop(k)
        else: os.environ[k] = v


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


print("configs:", [c[0] for c in configs], flush=True)
totals = {c[0]: [0.0] * 4 for c in configs}
for name, (n, t, h, w, ci, co, k, s, p) in LAYERS:
    g = K.make_geom(n, t, h, w, ci, co, k, s, p)
    stem = name == "stem"
    stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
    dy = torch.randn(n, g.To, g.Ho, g.Wo, g.Cout_p, device=dev).bfloat16()
    if stem:
        xs = torch.randn(n, t, h // 2, w // 2 + 3, 16, device=dev).bfloat16()
        ws = (torch.randn(g.Cout_p, k[0] * 4, 64, device=dev) / 20).bfloat16()
        ys = torch.empty(n, g.To, g.Ho, g.Wo, g.Cout_p, device=dev, dtype=torch.bfloat16)
        dws = torch.empty((g.Cout_p, 4, 64), dtype=torch.float32, device=dev)
        ops = [("fprop", lambda: _lib.call("dv_conv3d_stem_fprop_bf16", _lib.ptr(xs), _lib.ptr(ws), _lib.ptr(ys), _lib.ptr(stats),
                                           None, ctypes.byref(g), _lib.stream_ptr())),
               ("wgrad", lambda: _lib.call("dv_conv3d_stem_wgrad_bf16", _lib.ptr(xs), _lib.ptr(dy), _lib.ptr(dws), ctypes.byref(g),
                                           _lib.stream_ptr()))]
    else:
        x = torch.randn(n, t, h, w, g.Cin_p, device=dev).bfloat16()
        wt = torch.randn(co, ci, *k, device=dev) / 20
        use(configs[0])
        wf, wtt = K.pack_conv_weight(wt, g)
        ss = torch.randn(2 * g.Cin_p, device=dev)
        sums = torch.zeros(2 * g.Cin_p, dtype=torch.float64, device=dev)
        dxb = torch.empty_like(x)
        ops = [("fprop", lambda: K.conv3d_fprop(x, wf, g, bn_stats=stats)),
               ("dgrad", lambda: K.conv3d_dgrad(dy, wtt, g)),
               ("dgrad+bnred", lambda: _lib.call("dv_conv3d_dgrad_bnred_bf16", _lib.ptr(dy), _lib.ptr(wtt), _lib.ptr(dxb),
                                                 ctypes.byref(g), _lib.ptr(x), _lib.ptr(ss), _lib.ptr(sums), _lib.stream_ptr())),
               ("wgrad", lambda: K.conv3d_wgrad_packed(x, dy, g))]
    fl = 2.0 * n * g.To * g.Ho * g.Wo * co * ci * k[0] * k[1] * k[2] / 1e9
    res = {c[0]: [] for c in configs}
    for oi, (oname, fn) in enumerate(ops):
        ts = {c[0]: [] for c in configs}
        for cfg in configs:
            use(cfg); fn(); fn()
        torch.cuda.synchronize()
        for _ in range(REPS):
            for cfg in configs:
                use(cfg)
                ts[cfg[0]].append(timed(fn))
        for c in configs:
            med = sorted(ts[c[0]])[REPS // 2]
            res[c[0]].append((oname, med))
            totals[c[0]][{"fprop": 0, "dgrad": 1, "dgrad+bnred": 2, "wgrad": 3}[oname]] += med
    for c in configs:
        print(f"{name:20s} {c[0]:14s} " + " | ".join(f"{o} {m:6.3f} ms {fl/m:5.0f} TF/s" for o, m in res[c[0]]), flush=True)
for c in configs:
    print(f"{'sum':20s} {c[0]:14s} fprop {totals[c[0]][0]:6.3f} | dgrad {totals[c[0]][1]:6.3f} | dgrad+bnred {totals[c[0]][2]:6.3f} | wgrad {totals[c[0]][3]:6.3f}")
shutil.rmtree(tmpdir, ignore_errors=True)
