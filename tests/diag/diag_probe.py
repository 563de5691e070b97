import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
# the dv_debug_* entry points live in the diagnostics build only (`make diag`)
os.environ.setdefault("DV_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dualvar_b200", "lib", "libdualvar_b200_diag.so"))
import torch
from dualvar_b200 import _lib
src = torch.arange(8192, dtype=torch.int16, device="cuda")
out = torch.zeros(512, dtype=torch.int16, device="cuda")
try:
    _lib.call("dv_debug_probe_overlap_tmap", _lib.ptr(src), _lib.ptr(out), 5, _lib.stream_ptr())
    torch.cuda.synchronize()
    o = out.cpu().view(8, 8, 8)  # row, 16B chunk, elem
    for r in range(8):
        # un-swizzle: physical chunk c holds logical chunk c ^ r
        row = torch.cat([o[r, c ^ r] for c in range(8)])
        exp = torch.arange((5 + r) * 16, (5 + r) * 16 + 64, dtype=torch.int16)
        print(r, "ok" if torch.equal(row, exp) else "MISMATCH", row[:4].tolist(), row[-2:].tolist())
except Exception as e:
    print("FAILED:", e)
