"""Throughput of the frame-staging kernels (Scale + RandomCrop from decoded 320x240 frames) next to Pillow on the host."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from PIL import Image
from dualvar_b200 import frames as FR, engine as E

B, V, T, Hs, Ws = 64, 3, 16, 240, 320
frames = torch.randint(0, 256, (B, V * T, Hs, Ws, 3), dtype=torch.uint8, device="cuda")
crops = FR.draw_crops(B, V)
for _ in range(3):
    out = FR.scale_crop(frames, crops, V)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = FR.scale_crop(frames, crops, V)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
rd = frames.numel() + B * V * T * Hs * 128 * 4 + B * V * T * 160 * 112 * 4   # frames read, RGBX intermediate written, ~160 rows x 112 columns of it re-read
wr = out.numel()
print(f"GPU scale_crop: {B} samples x {V * T} frames {Ws}x{Hs}: {ms:.3f} ms/batch = {B / ms * 1e3:.0f} samples/s, "
      f"{(rd + wr) / ms / 1e6:.0f} GB/s of byte traffic", flush=True)
e0.record()
for _ in range(10):
    act = E.ingest(E.RawClips(out, V), s2d=True)
e1.record(); torch.cuda.synchronize()
print(f"GPU ingest of the staged clips (ToTensor + Normalize + space-to-depth bf16): {e0.elapsed_time(e1) / 10:.3f} ms/batch")
# the reference's way: PIL on the host, per frame (one core; the loader runs 16 workers)
f_np = frames[:2].cpu().numpy()
t0 = time.perf_counter()
for b in range(2):
    for f in range(V * T):
        im = Image.fromarray(f_np[b, f]).resize((128, 171), Image.BICUBIC)
        l, u = int(crops[b, f // T, 0]), int(crops[b, f // T, 1])
        np.asarray(im.crop((l, u, l + 112, u + 112)))
dt = (time.perf_counter() - t0) / 2
print(f"Pillow on one host core: {dt * 1e3:.1f} ms/sample = {1 / dt:.1f} samples/s/core ({os.cpu_count()} cores on this box)")
