"""Golden vectors for the frame-staging row: what the reference's loader computes for
A.Scale((128, 171)) -> A.RandomCrop(112) -> A.ToTensor() (utils/augmentation.py:125-176,361-364), produced with the
libraries the reference itself calls (Pillow's Image.resize(..., BICUBIC) / Image.crop and torchvision's ToTensor) and
Python's `random` in RandomCrop's draw order. Run in the build container:

    python tests/golden/make_golden_frames.py        # writes tests/golden/frames.npz (Pillow version recorded inside)
"""
import os
import random

import numpy as np
import PIL
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261018)
    B, V, T, Hs, Ws = 1, 3, 2, 96, 128
    frames = rng.integers(0, 256, (B, V * T, Hs, Ws, 3), dtype=np.uint8)
    # a hard-edged frame: bicubic over/undershoot must clip exactly like Pillow's lookup table
    frames[0, 0] = np.kron(rng.integers(0, 2, (Hs // 8, Ws // 8, 1), dtype=np.uint8) * 255, np.ones((8, 8, 3), np.uint8))
    random.seed(1018)
    crops = np.zeros((B, V, 2), np.int32)
    out = np.zeros((B, 3, V * T, 112, 112), np.uint8)
    tens = np.zeros((B, 3, V * T, 112, 112), np.float32)
    for b in range(B):
        for v in range(V):
            imgs = [Image.fromarray(frames[b, v * T + t]).resize((128, 171), Image.BICUBIC) for t in range(T)]
            h, w = imgs[0].size[0], imgs[0].size[1]                     # RandomCrop: "h" is PIL's width
            h_start = random.randint(0, h - 112)
            w_start = random.randint(0, w - 112)
            crops[b, v] = (h_start, w_start)
            for t, im in enumerate(imgs):
                c = im.crop((h_start, w_start, h_start + 112, w_start + 112))
                out[b, :, v * T + t] = np.asarray(c).transpose(2, 0, 1)
                tens[b, :, v * T + t] = transforms.ToTensor()(c).numpy()
    # a second geometry: down-scaling by 2.5 horizontally (11 taps), one frame
    big = rng.integers(0, 256, (120, 320, 3), dtype=np.uint8)
    big_resized = np.asarray(Image.fromarray(big).resize((128, 171), Image.BICUBIC))
    np.savez_compressed(os.path.join(HERE, "frames.npz"), frames=frames, crops=crops, out_u8=out,
                        out_tensor_checksum=np.array([float(tens.astype(np.float64).sum())]),
                        big=big, big_resized=big_resized, pillow_version=np.array(PIL.__version__), seed=np.array(1018))
    print("wrote frames.npz", out.shape, "Pillow", PIL.__version__)


if __name__ == "__main__":
    main()
