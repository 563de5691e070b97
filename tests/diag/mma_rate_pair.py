"""GPU diagnostic: tcgen05.mma.cta_group::2 issue rate at N = 64 (the issue-bound conv launches): one issuing thread
with the plain loop / with 12 straight-line MMAs + a commit per stage, and TWO issuing threads on separate
accumulators (whole stages each, or half of every stage each). Fetch floor at N = 64: 40 cycles per MMA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("DV_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dualvar_b200", "lib", "libdualvar_b200_diag.so"))
import torch
from dualvar_b200 import _lib
dev = "cuda:0"
out = torch.zeros(148 * 2, dtype=torch.int64, device=dev)
n_mma = 4800
names = {0x002: "1 thread, loop of 4", 0x102: "1 thread, 12 straight + commit", 0x202: "2 threads, whole stages each",
         0x302: "2 threads, half of every stage each"}
for n in (64, 128, 144, 256):
    for mode in (0x002, 0x102, 0x202, 0x302):
        for rep in range(2):
            _lib.call("dv_debug_mma_rate", n, n_mma, 160 * 1024, mode, _lib.ptr(out), 148, _lib.stream_ptr())
            torch.cuda.synchronize()
        c = out.view(148, 2)[0::2].double()
        issue, total = c[:, 0].mean().item() / n_mma, c[:, 1].mean().item() / n_mma
        print(f"N={n:3d} {names[mode]:38s}: issue {issue:6.1f} complete {total:6.1f} cyc/MMA (tensor floor {n/2:.0f}, "
              f"fetch floor {(4096 + 16 * n) / 128:.0f}) -> {n/2/total*100:4.0f}% of peak", flush=True)
