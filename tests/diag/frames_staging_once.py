"""Two staging calls + one ingest for an ncu launch list (tests/diag/frames_staging.py is the timed version)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dualvar_b200 import frames as FR, engine as E
B, V, T, Hs, Ws = 64, 3, 16, 240, 320
frames = torch.randint(0, 256, (B, V * T, Hs, Ws, 3), dtype=torch.uint8, device="cuda")
crops = FR.draw_crops(B, V)
for _ in range(2):
    out = FR.scale_crop(frames, crops, V)
act = E.ingest(E.RawClips(out, V), s2d=True)
torch.cuda.synchronize()
