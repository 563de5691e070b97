"""Golden vectors for the colour-jitter stage: the UNMODIFIED reference class ``A.ColorJitter``
(/root/reference/utils/augmentation.py:429-660), built as pretrain.py:505 builds it, run on seeded frames with seeded
``random`` / ``np.random``. Only runs in the build container (the reference does not travel):

    python tests/golden/make_golden_color_jitter.py      # writes tests/golden/color_jitter.npz
"""
import collections
import collections.abc
import importlib.util
import os
import random

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_augmentation():
    collections.Iterable = collections.abc.Iterable          # removed from `collections` in Python 3.10 (shim)
    spec = importlib.util.spec_from_file_location("ref_augmentation", "/root/reference/utils/augmentation.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    aug = load_reference_augmentation()
    rng = np.random.default_rng(77)
    out = {}
    for tag, consistent in (("per_frame", False), ("consistent", True)):
        u8 = rng.integers(0, 256, (8, 3, 24, 32), dtype=np.uint8)
        u8[0, :, :6, :6] = 128                                 # a grey patch: the max == min branch of the hue conversion
        frames = [torch.from_numpy(f).float().div(255) for f in u8]      # what A.ToTensor() hands to ColorJitter
        cj = aug.ColorJitter(0.8, 0.8, 0.8, 0.2, p=0.8, consistent=consistent, seq_len=4, block=1, grad_consistent=False)
        random.seed(2026); np.random.seed(1018)
        res = cj(frames)
        out[tag + "_u8"] = u8
        out[tag + "_out"] = torch.stack(res).numpy()
    np.savez_compressed(os.path.join(HERE, "color_jitter.npz"), py_seed=np.array(2026), np_seed=np.array(1018),
                        seq_len=np.array(4), **out)
    print("wrote color_jitter.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
