"""Timing of the colour-jitter kernel at the bench batch (64 samples x 48 frames of 112x112) next to torchvision on a host core."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from dualvar_b200 import frames as FR
B, F = 64, 48
clips = torch.randint(0, 256, (B, 3, F, 112, 112), dtype=torch.uint8, device="cuda")
random.seed(0); np.random.seed(0)
prm = FR.draw_color_jitter(B * F).cuda()
for _ in range(3):
    out = FR.color_jitter(clips, prm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = FR.color_jitter(clips, prm)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = clips.numel() + out.numel() * 4
print(f"GPU color_jitter: {B} samples x {F} frames 112x112 ({int(prm[:, 0].sum())} of {B * F} frames jittered): {ms:.3f} ms/batch = "
      f"{B / ms * 1e3:.0f} samples/s, {byt / ms / 1e6:.0f} GB/s (1 B read + 4 B written per value)", flush=True)
import torchvision.transforms.functional as TF
x = (clips[0, :, 0].cpu().float() / 255)
t0 = time.perf_counter()
for _ in range(20):
    y = TF.adjust_hue(TF.adjust_saturation(TF.adjust_contrast(TF.adjust_brightness(x, 1.2), 0.9), 1.3), 0.1)
dt = (time.perf_counter() - t0) / 20
print(f"torchvision on the host (default threads): {dt * 1e3:.2f} ms/frame = {dt * F * 1e3:.0f} ms/sample")
