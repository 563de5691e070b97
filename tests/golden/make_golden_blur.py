"""Golden vectors for the Gaussian-blur stage: the UNMODIFIED reference class ``A.GaussianBlur``
(/root/reference/utils/augmentation.py:706-721) on seeded frames with a seeded ``random``. Build container only:

    python tests/golden/make_golden_blur.py      # writes tests/golden/gaussian_blur.npz
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_color_jitter import load_reference_augmentation  # noqa: E402


def main():
    aug = load_reference_augmentation()
    rng = np.random.default_rng(99)
    u8 = rng.integers(0, 256, (8, 3, 24, 32), dtype=np.uint8)
    frames = [torch.from_numpy(f).float().div(255) for f in u8]
    gb = aug.GaussianBlur([.1, 2.], seq_len=4)
    random.seed(1234)
    res = torch.stack(gb(frames)).numpy()
    np.savez_compressed(os.path.join(HERE, "gaussian_blur.npz"), u8=u8, out=res, py_seed=np.array(1234), seq_len=np.array(4))
    print("wrote gaussian_blur.npz", res.shape)


if __name__ == "__main__":
    main()
