"""GPU diagnostic: max-pool forward / recorded-argmax backward time at the S3D-G sizes (64 clips = 16 samples x 4 passes)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dualvar_b200 import _lib
from dualvar_b200._lib import ptr, call, stream_ptr
dev = "cuda:0"
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (N, T, H, W, C, k, s, p) in [(64, 16, 16, 16, 192, (3, 3, 3), (1, 1, 1), (1, 1, 1)), (64, 8, 8, 8, 480, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
                                 (64, 16, 64, 64, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1)), (64, 16, 16, 16, 480, (3, 3, 3), (2, 2, 2), (1, 1, 1))]:
    To, Ho, Wo = [(d + 2 * pp - kk) // ss + 1 for d, kk, ss, pp in zip((T, H, W), k, s, p)]
    x = torch.relu(torch.randn(N, T, H, W, C, device=dev)).bfloat16()
    y = torch.empty(N, To, Ho, Wo, C, device=dev, dtype=torch.bfloat16)
    idx = torch.empty(N, To, Ho, Wo, C, device=dev, dtype=torch.uint8)
    dy = torch.randn_like(y); dx = torch.empty_like(x)
    geom = (ctypes.c_int32 * 17)(N, T, H, W, To, Ho, Wo, C, *k, *s, *p)
    f = timeit(lambda: call("dv_maxpool3d_fwd_idx", ptr(x), ptr(y), ptr(idx), geom, stream_ptr()))
    b = timeit(lambda: call("dv_maxpool3d_bwd_idx", ptr(idx), ptr(dy), ptr(dx), geom, stream_ptr()))
    gb = (x.numel() * 2 + y.numel() * 3) / 1e9
    print(f"{(N,T,H,W,C)} k{k} s{s}: fwd {f:7.1f} us ({gb/f*1e6:5.0f} GB/s), bwd {b:7.1f} us", flush=True)
