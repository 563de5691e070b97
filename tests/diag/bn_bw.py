"""GPU diagnostic: achieved HBM GB/s of the BatchNorm passes at the R(2+1)D layer-1 sizes (algorithmic bytes / time)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dualvar_b200 import _lib
from dualvar_b200._lib import ptr, stream_ptr, call
dev = "cuda:0"

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for rows, Cp in [(9633792, 144), (9633792, 128), (9633792, 160), (9633792, 64), (9633792, 88), (3211264, 144), (1204224, 288)]:
    y = torch.randn(rows, Cp, device=dev).bfloat16()
    dz = torch.randn(rows, Cp, device=dev).bfloat16()
    z = torch.empty_like(y); dy = torch.empty_like(y); gb = torch.empty_like(y)
    ss = torch.randn(2 * Cp, device=dev); coef = torch.randn(3 * Cp, device=dev)
    sums = torch.zeros(2 * Cp, dtype=torch.float64, device=dev)
    n = rows * Cp * 2 / 1e9
    cases = [
        ("bn_apply relu", 2 * n, lambda: call("dv_bn_apply", ptr(y), ptr(ss), None, None, None, ptr(z), rows, Cp, Cp, 0, 1, stream_ptr())),
        ("bn_apply relu +res", 3 * n, lambda: call("dv_bn_apply", ptr(y), ptr(ss), None, None, ptr(dz), ptr(z), rows, Cp, Cp, 0, 1, stream_ptr())),
        ("bn_apply relu +bn2", 3 * n, lambda: call("dv_bn_apply", ptr(y), ptr(ss), ptr(dz), ptr(ss), None, ptr(z), rows, Cp, Cp, 0, 1, stream_ptr())),
        ("bwd_reduce mask(ss)", 2 * n, lambda: call("dv_bn_bwd_reduce", ptr(dz), None, ptr(z), ptr(y), ptr(ss), ptr(sums), rows, Cp, Cp, 0, 1, stream_ptr())),
        ("bwd_reduce 2 grads mask(out)", 4 * n, lambda: call("dv_bn_bwd_reduce", ptr(dz), ptr(dy), ptr(z), ptr(y), None, ptr(sums), rows, Cp, Cp, 0, 1, stream_ptr())),
        ("bwd_apply mask(ss)", 3 * n, lambda: call("dv_bn_bwd_apply", ptr(dz), None, ptr(z), ptr(y), ptr(ss), ptr(coef), ptr(dy), None, rows, Cp, Cp, 0, 1, stream_ptr())),
        ("bwd_apply 2 grads mask(out) +g", 6 * n, lambda: call("dv_bn_bwd_apply", ptr(dz), ptr(dy), ptr(z), ptr(y), None, ptr(coef), ptr(dy), ptr(gb), rows, Cp, Cp, 0, 1, stream_ptr())),
        ("torch copy (yardstick)", 2 * n, lambda: z.copy_(y)),
    ]
    for name, gbytes, fn in cases:
        ms = timeit(fn)
        print(f"rows {rows:8d} Cp {Cp:3d} {name:32s} {ms:7.3f} ms  {gbytes/ms:7.1f} GB/s", flush=True)
