"""Colour jitter (second stage): A.ColorJitter as pretrain.py:505 builds it.

CPU: the oracle's four operations are torchvision's tensor ops bit for bit; oracle + RNG mirror reproduce the UNMODIFIED
reference class (tests/golden/color_jitter.npz, written by tests/golden/make_golden_color_jitter.py) bit for bit; the
product's draw_color_jitter equals the oracle's.
GPU: dv_frames_color_jitter through the C ABI against the golden vectors and the oracle. Tolerance 2e-6 absolute on
[0, 1] values: every pixel operation follows torchvision's rounding sequence, only adjust_contrast's frame mean is summed
in a different order (unjittered frames must be bit-exact).
"""
import os
import random

import numpy as np
import pytest
import torch

ATOL = 2e-6


def test_oracle_ops_are_torchvision_bit_for_bit():
    import torchvision.transforms.functional as F
    from oracle import augment as A
    torch.manual_seed(0)
    img = torch.randint(0, 256, (3, 40, 56), dtype=torch.uint8).float() / 255
    img[:, :5, :5] = 0.5
    for mine, ref, fs in ((A.adjust_brightness, F.adjust_brightness, (0.2, 1.0, 1.8)),
                          (A.adjust_contrast, F.adjust_contrast, (0.2, 1.3, 1.8)),
                          (A.adjust_saturation, F.adjust_saturation, (0.2, 1.3, 1.8)),
                          (A.adjust_hue, F.adjust_hue, (-0.2, -0.05, 0.0, 0.13, 0.2))):
        for f in fs:
            assert torch.equal(mine(img, f), ref(img, f)), (ref.__name__, f)


@pytest.mark.parametrize("tag,consistent", [("per_frame", False), ("consistent", True)])
def test_oracle_and_rng_mirror_reproduce_the_reference_class(golden_dir, tag, consistent):
    from dualvar_b200 import frames as FR
    from oracle import augment as A
    g = np.load(os.path.join(golden_dir, "color_jitter.npz"))
    random.seed(int(g["py_seed"])); np.random.seed(int(g["np_seed"]))
    prm = A.draw_color_jitter(8, random, np.random, consistent=consistent, seq_len=int(g["seq_len"]))
    random.seed(int(g["py_seed"])); np.random.seed(int(g["np_seed"]))
    prm_product = FR.draw_color_jitter(8, consistent=consistent, seq_len=int(g["seq_len"]))
    assert np.array_equal(prm, prm_product.numpy())
    assert 0 < prm[:, 0].sum() < 8 or consistent              # the golden has jittered and unjittered frames
    out = A.color_jitter(torch.from_numpy(g[tag + "_u8"]).float().div(255), prm)
    assert torch.equal(out, torch.from_numpy(g[tag + "_out"]))


def test_color_jitter_refuses_cpu_tensors():
    from dualvar_b200 import _lib, frames as FR
    with pytest.raises(_lib.DualVarNativeError):
        FR.color_jitter(torch.zeros((1, 3, 2, 8, 8), dtype=torch.uint8), torch.zeros((2, 12)))


@pytest.mark.gpu
@pytest.mark.parametrize("tag,consistent", [("per_frame", False), ("consistent", True)])
def test_gpu_color_jitter_matches_reference_golden(golden_dir, tag, consistent):
    from dualvar_b200 import frames as FR
    g = np.load(os.path.join(golden_dir, "color_jitter.npz"))
    random.seed(int(g["py_seed"])); np.random.seed(int(g["np_seed"]))
    prm = FR.draw_color_jitter(8, consistent=consistent, seq_len=int(g["seq_len"]))
    u8 = torch.from_numpy(g[tag + "_u8"])                                   # (8 frames, 3, 24, 32)
    clips = u8.permute(1, 0, 2, 3).unsqueeze(0).contiguous().cuda()           # one sample of 8 frames: (1, 3, 8, 24, 32)
    got = FR.color_jitter(clips, prm).cpu()[0].permute(1, 0, 2, 3)
    want = torch.from_numpy(g[tag + "_out"])
    err = (got - want).abs().amax(dim=(1, 2, 3))
    print(tag, "max abs error per frame", [f"{e:.1e}" for e in err.tolist()])
    assert float(err.max()) <= ATOL
    for i in range(8):
        if prm[i, 0] == 0:
            assert torch.equal(got[i], want[i])                             # ToTensor only: bit-exact


@pytest.mark.gpu
def test_gpu_color_jitter_at_crop_size_and_through_stage_clips():
    from dualvar_b200 import engine as E, frames as FR
    from oracle import augment as A
    from oracle.frames import scale_crop
    rng = np.random.default_rng(3)
    B, V, T = 2, 3, 2
    dec = rng.integers(0, 256, (B, V * T, 120, 160, 3), dtype=np.uint8)
    dec[0, 0, :40, :40] = 77                                                 # grey region: max == min branch of the hue op
    random.seed(5); np.random.seed(6)
    crops = FR.draw_crops(B, V)
    prm = FR.draw_color_jitter(B * V * T)
    want_u8 = torch.from_numpy(scale_crop(dec, crops.numpy(), V))            # (B, 3, F, 112, 112)
    frames = want_u8.permute(0, 2, 1, 3, 4).reshape(B * V * T, 3, 112, 112).float().div(255)
    want = A.color_jitter(frames, prm.numpy()).view(B, V * T, 3, 112, 112).permute(0, 2, 1, 3, 4)
    got = FR.color_jitter(torch.from_numpy(scale_crop(dec, crops.numpy(), V)).cuda(), prm).cpu()
    assert float((got - want).abs().max()) <= ATOL
    # end to end: decoded frames -> Scale/RandomCrop -> ColorJitter -> Normalize -> bf16 NDHWC
    act = E.ingest(FR.stage_clips(torch.from_numpy(dec).cuda(), crops, V, jitter=prm))
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)
    x = ((want - mean) * (1.0 / std)).view(B, 3, V, T, 112, 112).permute(0, 2, 3, 4, 5, 1).reshape(B * V, T, 112, 112, 3)
    d = (act.data[..., :3].float().cpu() - x).abs().max().item()
    assert d <= 0.02                                                         # bf16 rounding of values up to ~2.7
