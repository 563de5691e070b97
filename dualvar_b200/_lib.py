"""ctypes binding of the C-ABI library (include/dualvar_b200.h).

The product path has no CPU or eager fallback: if the shared library is missing, importing any
compute entry point raises. Build it with ``make`` (or ``python -c "import __graft_entry__ as g;
g.build()"``) at the repo root; the built ``dualvar_b200/lib/libdualvar_b200.so`` stays in-tree.
"""
import ctypes
import os
import re

import torch

# DV_LIB_PATH: load another build of the same library (A/B runs of kernel changes in one GPU call)
_LIB_PATH = os.environ.get("DV_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                          "libdualvar_b200.so")
_lib = None


class DualVarNativeError(RuntimeError):
    pass


class ConvGeom(ctypes.Structure):
    """Mirror of ``dv_conv_geom``."""

    _fields_ = [(n, ctypes.c_int32) for n in (
        "N", "T", "H", "W", "Cin", "Cout", "Cin_p", "Cout_p",
        "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw", "To", "Ho", "Wo")]

    @property
    def taps(self):
        return self.kt * self.kh * self.kw

    def out_positions(self):
        return self.N * self.To * self.Ho * self.Wo

    def in_positions(self):
        return self.N * self.T * self.H * self.W


def pad8(c):
    return (int(c) + 7) // 8 * 8


def make_geom(N, T, H, W, Cin, Cout, kernel, stride, padding):
    kt, kh, kw = kernel
    st, sh, sw = stride
    pt, ph, pw = padding
    g = ConvGeom()
    g.N, g.T, g.H, g.W = N, T, H, W
    g.Cin, g.Cout, g.Cin_p, g.Cout_p = Cin, Cout, pad8(Cin), pad8(Cout)
    g.kt, g.kh, g.kw = kt, kh, kw
    g.st, g.sh, g.sw = st, sh, sw
    g.pt, g.ph, g.pw = pt, ph, pw
    g.To = (T + 2 * pt - kt) // st + 1
    g.Ho = (H + 2 * ph - kh) // sh + 1
    g.Wo = (W + 2 * pw - kw) // sw + 1
    return g


_P = ctypes.c_void_p
_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dualvar_b200.h")
_SCALARS = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64,
            "float": ctypes.c_float, "double": ctypes.c_double}


def _parse_header(path=_HEADER):
    """Derive the ctypes signatures from the public header so the binding cannot drift from it."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    sigs = {}
    for m in re.finditer(r"\b(int64_t|int|const char\*)\s+(dv_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(_P)
                else:
                    base = a.replace("const ", "").split(" ")[0]
                    argtypes.append(_SCALARS[base])
        sigs[name] = (ctypes.c_char_p if "char" in ret else (ctypes.c_int64 if ret == "int64_t" else ctypes.c_int), argtypes)
    return sigs


_SIGNATURES = _parse_header()
# DV_LIB_PATH pointing at the diagnostics build (libdualvar_b200_diag.so, `make diag`): the dv_debug_* entry points of
# include/dualvar_b200_diag.h are bound as well. The product library has none of them.
_DIAG_HEADER = os.path.join(os.path.dirname(_HEADER), "dualvar_b200_diag.h")
if "_diag" in os.path.basename(_LIB_PATH) and os.path.exists(_DIAG_HEADER):
    _SIGNATURES.update(_parse_header(_DIAG_HEADER))


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise DualVarNativeError(
            f"{_LIB_PATH} is missing: the CUDA extension is not built (run `make` at the repo "
            "root). dualvar_b200 has no CPU/eager fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            # an OLDER build loaded through DV_LIB_PATH for an A/B run may lack entry points this header declares;
            # the product library itself must export every one of them (tests/test_abi.py)
            if os.environ.get("DV_LIB_PATH"):
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (raw query: no Stream object per launch)."""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def ptr(t):
    """Device address of a tensor as a plain int (ctypes converts it for the void* parameters), None for NULL."""
    return None if t is None else t.data_ptr()


def check(rc, what):
    if rc != 0:
        msg = load().dv_last_error()
        raise DualVarNativeError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


_TRACE = os.environ.get("DV_TRACE", "") not in ("", "0")
launch_count = 0


def _fmt(a):
    if isinstance(a, ConvGeom):
        return "geom(" + ",".join(f"{n}={getattr(a, n)}" for n, _ in ConvGeom._fields_) + ")"
    if hasattr(a, "_obj"):
        return _fmt(a._obj)
    if isinstance(a, ctypes.c_void_p):
        return hex(a.value or 0)
    if hasattr(a, "value"):
        return str(a.value)
    return str(a)


class KernelTimer:
    """CUDA-event timing of selected C-ABI calls on the launching stream (bench.py roofline leg).
    For conv calls the algorithmic FLOPs (2 * positions * Cout * Cin * taps, logical channels) are
    accumulated alongside so that achieved TFLOP/s = flops / device time."""

    def __init__(self, names, detail=False):
        self.names = set(names)
        self.detail = detail
        self.records = []   # (name, start_event, end_event, flops, bytes)

    def key_of(self, name, args):
        if not self.detail:
            return name
        for a in args:
            g = getattr(a, "_obj", None)
            if isinstance(g, ConvGeom):
                return (f"{name} N{g.N} {g.T}x{g.H}x{g.W} {g.Cin}->{g.Cout} k{g.kt}{g.kh}{g.kw} "
                        f"s{g.st}{g.sh}{g.sw}")
        pos = {"dv_bn_apply": (6, 7), "dv_bn_bwd_reduce": (6, 7), "dv_bn_bwd_apply": (8, 9)}.get(name)
        if pos is not None:      # rows x padded channels + which optional streams are present
            opt = {"dv_bn_apply": (2, 4), "dv_bn_bwd_reduce": (1,), "dv_bn_bwd_apply": (1, 7)}[name]
            flags = "".join("1" if args[i] is not None else "0" for i in opt)
            return f"{name} rows{args[pos[0]]} C{args[pos[1]]} opt{flags}"
        ints = [str(a) for a in args if isinstance(a, int) and 0 <= a < (1 << 32)][:3]
        return name + " " + ",".join(ints)

    def flops_of(self, args):
        for a in args:
            g = getattr(a, "_obj", None)
            if isinstance(g, ConvGeom):
                return 2.0 * g.N * g.To * g.Ho * g.Wo * g.Cout * g.Cin * g.taps
        return 0.0

    def bytes_of(self, name, args):
        """Algorithmic HBM bytes of one call on LOGICAL channel counts: every operand read once, every result written
        once (bf16 activations / gradients 2 B, packed bf16 weights 2 B, fp32 weight gradients 4 B).
        BatchNorm passes (rows, Cp follow the pointer arguments): apply reads y and writes z (+ second branch /
        residual reads), bwd_reduce reads dz and y (+ dz2 / out), bwd_apply reads dz and y and writes dy (+ g)."""
        for a in args:
            g = getattr(a, "_obj", None)
            if isinstance(g, ConvGeom):
                xin = g.N * g.T * g.H * g.W * g.Cin
                yout = g.N * g.To * g.Ho * g.Wo * g.Cout
                w = g.Cout * g.Cin * g.taps
                if "wgrad" in name:
                    return 2.0 * (xin + yout) + 4.0 * w
                return 2.0 * (xin + yout) + 2.0 * w + (2.0 * xin if "bnred" in name else 0.0)
        def n(*xs):
            return sum(1 for x in xs if x is not None)
        if name == "dv_bn_apply":            # (y1, ss1, y2, ss2, res, out, rows, Cp, ...)
            return 2.0 * args[6] * args[7] * (n(args[0], args[2], args[4]) + 1)
        if name == "dv_bn_bwd_reduce":       # (dout, dout2, out, y, mask_ss, sums, rows, Cp, o_ld, o_coff, relu, stream)
            reads_out = args[4] is None and args[10] != 0
            return 2.0 * args[6] * args[7] * (n(args[0], args[1], args[3]) + (1 if reads_out else 0))
        if name == "dv_bn_bwd_apply":        # (dout, dout2, out, y, mask_ss, coef, dy, g_out, rows, Cp, o_ld, o_coff, relu, stream)
            reads_out = args[4] is None and args[12] != 0
            return 2.0 * args[8] * args[9] * (n(args[0], args[1], args[3]) + (1 if reads_out else 0) + n(args[6], args[7]))
        return 0.0

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, fl, by in self.records:
            d = out.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += by
        return out


_timer = None
_NVTX = KernelTimer([], detail=True) if os.environ.get("DV_NVTX", "") not in ("", "0") else None


def set_timer(timer):
    global _timer
    _timer = timer


_fn_cache = {}


def call(name, *args):
    """Invoke a C-ABI function that returns a status code; raise on failure."""
    global launch_count
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    launch_count += 1
    if _timer is not None and name in _timer.names:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        check(rc, name)
        _timer.records.append((_timer.key_of(name, args), e0, e1, _timer.flops_of(args), _timer.bytes_of(name, args)))
        return
    if _TRACE:
        print(f"[dv] {name}(" + ", ".join(_fmt(a) for a in args) + ")", flush=True)
    if _NVTX is not None:
        # DV_NVTX=1: every C-ABI call sits in an NVTX range named like KernelTimer's per-layer key, so that
        # `ncu --nvtx --print-nvtx-rename kernel` reports launches per layer (tests/diag/ncu_step.py)
        torch.cuda.nvtx.range_push(_NVTX.key_of(name, args).replace(" ", "_"))
        rc = fn(*args)
        torch.cuda.nvtx.range_pop()
        check(rc, name)
        return
    rc = fn(*args)
    if rc != 0:
        check(rc, name)
    if _TRACE:
        torch.cuda.synchronize()
        print(f"[dv] {name} done", flush=True)
