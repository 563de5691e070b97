"""GPU diagnostic: steady-state tcgen05.mma rate (SS mode, M=128, K=16) vs N — the operand-fetch ceiling the conv
kernels live under. Ideal = N/2 cycles per MMA (8192 MAC/clk/SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
# the dv_debug_* entry points live in the diagnostics build only (`make diag`)
os.environ.setdefault("DV_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dualvar_b200", "lib", "libdualvar_b200_diag.so"))
import torch
from dualvar_b200 import _lib
dev = "cuda:0"
out = torch.zeros(148 * 2, dtype=torch.int64, device=dev)
n_mma = 4096
for grid in (148,):
    # 0x2pp: commit to a rotating mbarrier every pp MMAs; 0x3pp: and switch accumulator at every commit
    # 3: straight-line issue with precomputed descriptors
    for mode in (0, 3, 0x20C, 0x224):
        for n in (64, 144, 256):
            for region in (160 * 1024,):
                _lib.call("dv_debug_mma_rate", n, n_mma, region, mode, _lib.ptr(out), grid, _lib.stream_ptr())
                torch.cuda.synchronize()
                c = out.view(148, 2)[:grid].double()
                issue, total = c[:, 0].mean().item() / n_mma, c[:, 1].mean().item() / n_mma
                bytes_per = 4096 + n * 32 // (2 if mode == 2 else 1)
                if mode == 2:
                    c = c[0::2]      # leader CTAs
                    issue, total = c[:, 0].mean().item() / n_mma, c[:, 1].mean().item() / n_mma
                print(f"grid {grid:3d} mode {mode:#x} N={n:3d} region {region//1024:3d}K: issue {issue:6.1f} cyc/MMA, "
                      f"complete {total:6.1f} cyc/MMA (ideal {n/2:.0f}) -> {bytes_per/total:5.1f} B/clk operand fetch, "
                      f"{n/2/total*100:4.0f}% of peak", flush=True)
