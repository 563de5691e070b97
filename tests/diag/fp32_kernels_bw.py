"""GPU diagnostic: the fp32-mode kernels at R(2+1)D layer-1 sizes (B=16: 48 clip-passes): achieved HBM GB/s of the
fp32 BatchNorm passes (algorithmic bytes / time) and TFLOP/s of one plane product / of a whole fp32 convolution."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dualvar_b200 import _lib, engine as E
from dualvar_b200._lib import ptr, stream_ptr, call, make_geom
dev = "cuda:0"
K = 3


def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rows, Cp in [(48 * 16 * 56 * 56, 144), (48 * 16 * 56 * 56, 64)]:
    y = torch.randn(rows, Cp, device=dev)
    dz = torch.randn(rows, Cp, device=dev)
    z = torch.empty_like(y)
    planes = torch.empty((K, rows, Cp), dtype=torch.bfloat16, device=dev)
    ss = torch.randn(2 * Cp, device=dev); coef = torch.randn(3 * Cp, device=dev)
    sums = torch.zeros(2 * Cp, dtype=torch.float64, device=dev)
    e = rows * Cp / 1e9            # G elements
    cases = [
        ("f32_colstats", 4 * e, lambda: call("dv_f32_colstats", ptr(y), ptr(sums), rows, Cp, stream_ptr())),
        ("f32_bn_apply relu (+3 planes)", (4 + 4 + 2 * K) * e, lambda: call("dv_f32_bn_apply", ptr(y), ptr(ss), None, None, None, ptr(z), ptr(planes), planes.stride(0), K, rows, Cp, Cp, 0, 1, stream_ptr())),
        ("f32_bn_bwd_reduce mask(ss)", 8 * e, lambda: call("dv_f32_bn_bwd_reduce", ptr(dz), None, None, ptr(y), ptr(ss), ptr(sums), rows, Cp, Cp, 0, 1, stream_ptr())),
        ("f32_bn_bwd_apply mask(ss) (3 planes out)", (8 + 2 * K) * e, lambda: call("dv_f32_bn_bwd_apply", ptr(dz), None, None, ptr(y), ptr(ss), ptr(coef), ptr(planes), planes.stride(0), K, None, rows, Cp, Cp, 0, 1, stream_ptr())),
        ("torch copy fp32 (yardstick)", 8 * e, lambda: z.copy_(y)),
    ]
    for name, gbytes, fn in cases:
        ms = timeit(fn)
        print(f"rows {rows:8d} Cp {Cp:3d} {name:42s} {ms:7.3f} ms  {gbytes / ms * 1e3:7.1f} GB/s", flush=True)
    del y, dz, z, planes

# one plane product and a whole 6-product convolution: 64 -> 144, (1,3,3), 48 x 16 x 56 x 56
E.set_precision("fp32", 3)
N, T, H, W, Cin, Cout = 48, 16, 56, 56, 64, 144
g = make_geom(N, T, H, W, Cin, Cout, (1, 3, 3), (1, 1, 1), (0, 1, 1))
conv = torch.nn.Conv3d(Cin, Cout, (1, 3, 3), 1, (0, 1, 1), bias=False).to(dev)
wp = E.packed_weight_planes(conv)
xp = torch.randn(K, N, T, H, W, Cin, device=dev).bfloat16()
yf = torch.zeros(N, T, H, W, g.Cout_p, device=dev)
yb = torch.empty(N, T, H, W, g.Cout_p, dtype=torch.bfloat16, device=dev)
wf_bf16, _ = E.packed_weights(conv)
fl = 2.0 * N * T * H * W * Cout * Cin * 9
ms1 = timeit(lambda: call("dv_conv3d_fprop_f32acc", ptr(xp[0]), ptr(wp[0][0]), ptr(yf), None, ctypes.byref(g), 1, stream_ptr()))
ms0 = timeit(lambda: call("dv_conv3d_fprop_bf16", ptr(xp[0]), ptr(wf_bf16), ptr(yb), None, None, ctypes.byref(g), stream_ptr()))


def whole():
    for n, (i, j) in enumerate(E._terms()):
        call("dv_conv3d_fprop_f32acc", ptr(xp[i]), ptr(wp[j][0]), ptr(yf), None, ctypes.byref(g), 1 if n else 0, stream_ptr())


ms6 = timeit(whole)
print(f"fprop 64->144 3x3 N48: bf16 kernel (bf16 TMA-store epilogue) {ms0:.3f} ms = {fl / ms0 / 1e9:.0f} TFLOP/s; one fp32-accumulate plane "
      f"product {ms1:.3f} ms = {fl / ms1 / 1e9:.0f} TFLOP/s of bf16 work; whole fp32 convolution (6 products, the first one storing) "
      f"{ms6:.3f} ms = {fl / ms6 / 1e9:.0f} TFLOP/s of fp32-equivalent work")
