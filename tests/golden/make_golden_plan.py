"""Golden vectors for the whole loader transform: the UNMODIFIED reference classes composed exactly as
pretrain.py:491-529 (get_transform) composes them - Scale, RandomCrop, ToTensor, RandomApply(ColorJitter, 0.8),
RandomApply(GaussianBlur, 0.5) in three branches under MultiRandomizedTransform with the same weights - at a reduced
geometry (Scale (32, 43), crop 28, 2 frames per clip) so that the file stays small. Build container only:

    python tests/golden/make_golden_plan.py      # writes tests/golden/transform_plan.npz
"""
import os
import random
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_color_jitter import load_reference_augmentation  # noqa: E402

SEQ, SCALED, CROP = 2, (32, 43), 28


def build_transform(A):
    def branch(jittered):
        tl = [A.Scale(SCALED), A.RandomCrop(size=CROP), A.ToTensor()]
        if jittered:
            tl += [transforms.RandomApply([A.ColorJitter(0.8, 0.8, 0.8, 0.2, p=0.8, consistent=False, seq_len=SEQ, block=1,
                                                          grad_consistent=False)], p=0.8),
                   transforms.RandomApply([A.GaussianBlur([.1, 2.], seq_len=SEQ)], p=0.5)]
        return transforms.Compose(tl)
    weights = [[0.2, 0.8, 0], [0, 1.0, 0], [0, 0., 1.0]]
    return A.MultiRandomizedTransform([branch(False), branch(True), branch(True)], SEQ, weights=weights)


def main():
    A = load_reference_augmentation()
    tr = build_transform(A)
    rng = np.random.default_rng(4)
    n = 6
    frames = rng.integers(0, 256, (n, 3 * SEQ, 40, 56, 3), dtype=np.uint8)
    random.seed(11); np.random.seed(12); torch.manual_seed(13)
    outs = []
    for b in range(n):
        seq = [Image.fromarray(f) for f in frames[b]]
        outs.append(torch.stack(tr(seq), dim=1).numpy())               # (3, 3*SEQ, 28, 28) as dataset/local_dataset.py:300
    np.savez_compressed(os.path.join(HERE, "transform_plan.npz"), frames=frames, out=np.stack(outs),
                        seeds=np.array([11, 12, 13]), seq_len=np.array(SEQ), scaled=np.array(SCALED), crop=np.array(CROP))
    print("wrote transform_plan.npz", np.stack(outs).shape)


if __name__ == "__main__":
    main()
