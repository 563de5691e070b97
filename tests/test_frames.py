"""Frame staging (first stage): Scale((128,171)) bicubic + RandomCrop(112) + ToTensor.

CPU: the numpy oracle (oracle/frames.py) is pinned bit-for-bit against Pillow itself - the third-party library the
reference calls (utils/augmentation.py:131-147) - and the library's host-side fixed-point tables against the oracle.
GPU: dv_frames_scale_crop_u8 through the C ABI against the oracle and Pillow, bit-exact, and the staged clips through
the ingest kernel against ToTensor + Normalize on the oracle's frames.
"""
import random

import numpy as np
import pytest
import torch

SIZES = [((240, 320), (128, 171)), ((256, 340), (128, 171)), ((80, 100), (128, 171)), ((171, 128), (128, 171)),
         ((239, 317), (128, 171)), ((240, 320), (171, 128)), ((60, 320), (128, 171))]


@pytest.mark.parametrize("src,dst", SIZES)
def test_oracle_resize_is_pillow_bit_for_bit(src, dst):
    from PIL import Image
    from oracle.frames import pil_bicubic_resize
    rng = np.random.default_rng(src[0] * 1000 + src[1])
    for img in (rng.integers(0, 256, src + (3,), dtype=np.uint8),
                np.kron(rng.integers(0, 2, (src[0] // 8 + 1, src[1] // 8 + 1, 1), dtype=np.uint8) * 255,
                        np.ones((8, 8, 3), np.uint8))[:src[0], :src[1]]):       # hard edges: over/undershoot clipping
        ref = np.asarray(Image.fromarray(img).resize(dst, Image.BICUBIC))
        assert np.array_equal(pil_bicubic_resize(img, *dst), ref)


def test_oracle_scale_crop_matches_reference_transform_chain():
    """oracle.scale_crop == A.Scale((128,171)) -> img.crop((h_start, w_start, ...)) -> ToTensor, written with PIL and
    torchvision exactly as utils/augmentation.py:131-176,361-364 does, including the RNG order of RandomCrop."""
    from PIL import Image
    from torchvision import transforms
    from oracle.frames import draw_crops, scale_crop
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (2, 6, 90, 120, 3), dtype=np.uint8)        # 2 samples x 3 views x 2 frames
    random.seed(3)
    crops = draw_crops(2, 3, random)
    got = scale_crop(frames, crops, 3)
    random.seed(3)
    for b in range(2):
        for v in range(3):
            imgs = [Image.fromarray(frames[b, v * 2 + t]).resize((128, 171), Image.BICUBIC) for t in range(2)]
            h, w = imgs[0].size[0], imgs[0].size[1]
            h_start, w_start = random.randint(0, h - 112), random.randint(0, w - 112)
            for t, im in enumerate(imgs):
                ten = transforms.ToTensor()(im.crop((h_start, w_start, h_start + 112, w_start + 112)))
                assert torch.equal(ten, torch.from_numpy(got[b, :, v * 2 + t]).float() / 255)


@pytest.mark.parametrize("n_in,n_out", [(320, 128), (240, 171), (100, 128), (80, 171), (128, 128), (317, 128), (1920, 128)])
def test_library_host_tables_match_oracle(n_in, n_out):
    """No GPU needed: the C++ restatement of Pillow's precompute_coeffs / normalize_coeffs_8bpc behind
    dv_frames_axis_table_host gives the oracle's tables integer for integer."""
    from dualvar_b200 import frames as FR
    from oracle.frames import resample_coeffs
    ks, t = FR.axis_table(n_in, n_out)
    xmin, xmax, kk = resample_coeffs(n_in, n_out)
    assert ks == kk.shape[1]
    assert np.array_equal(t[:, 0].numpy(), xmin) and np.array_equal(t[:, 1].numpy(), xmax)
    assert np.array_equal(t[:, 2:].numpy(), kk)


def test_draw_crops_mirrors_oracle_rng_order():
    from dualvar_b200 import frames as FR
    from oracle.frames import draw_crops
    random.seed(9); a = FR.draw_crops(5, 3)
    random.seed(9); b = draw_crops(5, 3, random)
    assert np.array_equal(a.numpy(), b)
    assert a[..., 0].max() <= 16 and a[..., 1].max() <= 59 and a.min() >= 0


def test_scale_crop_refuses_cpu_tensors():
    from dualvar_b200 import _lib, frames as FR
    with pytest.raises(_lib.DualVarNativeError):
        FR.scale_crop(torch.zeros((1, 3, 8, 8, 3), dtype=torch.uint8), torch.zeros((1, 3, 2), dtype=torch.int32), 3)


@pytest.mark.gpu
@pytest.mark.parametrize("src", [(240, 320), (256, 340), (80, 100), (171, 128), (239, 317), (60, 320),
                                 (480, 200), (700, 600)])       # 13-tap vertical windows; > 16 taps: the generic kernels
def test_gpu_scale_crop_bit_exact(src):
    from PIL import Image
    from dualvar_b200 import frames as FR
    from oracle.frames import scale_crop
    rng = np.random.default_rng(src[0])
    B, V, T = 2, 3, 4
    frames = rng.integers(0, 256, (B, V * T) + src + (3,), dtype=np.uint8)
    frames[0, 0] = np.kron(rng.integers(0, 2, (src[0] // 4 + 1, src[1] // 4 + 1, 1), dtype=np.uint8) * 255,
                           np.ones((4, 4, 3), np.uint8))[:src[0], :src[1]]
    random.seed(src[1])
    crops = FR.draw_crops(B, V)
    crops[0, 0] = torch.tensor([0, 0]); crops[1, 2] = torch.tensor([16, 59])           # the extreme windows
    got = FR.scale_crop(torch.from_numpy(frames).cuda(), crops, V).cpu().numpy()
    want = scale_crop(frames, crops.numpy(), V)
    assert got.shape == want.shape == (B, 3, V * T, 112, 112)
    assert np.array_equal(got, want)
    # and straight against Pillow for one frame
    b, f = 1, 2 * T + 1
    r = np.asarray(Image.fromarray(frames[b, f]).resize((128, 171), Image.BICUBIC))
    left, upper = int(crops[b, 2, 0]), int(crops[b, 2, 1])
    assert np.array_equal(got[b, :, f], r[upper:upper + 112, left:left + 112].transpose(2, 0, 1))


@pytest.mark.gpu
def test_gpu_staged_clips_feed_the_encoder_like_the_reference_loader():
    """stage_clips -> ingest == Normalize(ToTensor(crop(Scale(frame)))) of pretrain.py:386-389 on the oracle's frames
    (bf16 NDHWC, bit-exact), and a model step runs on the staged clips."""
    from dualvar_b200 import engine as E, frames as FR
    from oracle.frames import scale_crop
    rng = np.random.default_rng(1)
    B, V, T = 2, 3, 8
    frames = rng.integers(0, 256, (B, V * T, 120, 160, 3), dtype=np.uint8)
    random.seed(4)
    crops = FR.draw_crops(B, V)
    clips = FR.stage_clips(torch.from_numpy(frames).cuda(), crops, V)
    act = E.ingest(clips)
    want_u8 = torch.from_numpy(scale_crop(frames, crops.numpy(), V)).cuda()             # (B, 3, V*T, 112, 112)
    x = want_u8.float() / 255
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1, 1)
    x = ((x - mean) * (1.0 / std)).view(B, 3, V, T, 112, 112).permute(0, 2, 3, 4, 5, 1).reshape(B * V, T, 112, 112, 3)
    assert torch.equal(act.data[..., :3], x.bfloat16())
    assert bool((act.data[..., 3:] == 0).all())


def test_oracle_matches_committed_pillow_golden(golden_dir):
    """tests/golden/frames.npz was written by tests/golden/make_golden_frames.py with Pillow / torchvision (the libraries
    the reference's transform chain calls) and RandomCrop's own draw order; the oracle reproduces it without them."""
    import os
    from oracle.frames import draw_crops, pil_bicubic_resize, scale_crop
    g = np.load(os.path.join(golden_dir, "frames.npz"))
    random.seed(int(g["seed"]))
    crops = draw_crops(g["frames"].shape[0], 3, random)
    assert np.array_equal(crops, g["crops"])
    out = scale_crop(g["frames"], crops, 3)
    assert np.array_equal(out, g["out_u8"])
    assert abs(float((out.astype(np.float32) / 255).astype(np.float64).sum()) - float(g["out_tensor_checksum"][0])) < 1e-6
    assert np.array_equal(pil_bicubic_resize(g["big"], 128, 171), g["big_resized"])


@pytest.mark.gpu
def test_gpu_scale_crop_matches_committed_pillow_golden(golden_dir):
    import os
    from dualvar_b200 import frames as FR
    g = np.load(os.path.join(golden_dir, "frames.npz"))
    got = FR.scale_crop(torch.from_numpy(g["frames"]).cuda(), torch.from_numpy(g["crops"]), 3).cpu().numpy()
    assert np.array_equal(got, g["out_u8"])
    # the 320-wide frame (11-tap horizontal windows) through a full-height "crop": scale 120 -> 171 rows, crop all of it
    big = torch.from_numpy(g["big"]).cuda()[None, None]
    full = FR.scale_crop(big, torch.zeros((1, 1, 2), dtype=torch.int32), 1, crop_size=(128, 171)).cpu().numpy()
    assert np.array_equal(full[0, :, 0].transpose(1, 2, 0), g["big_resized"])


def test_center_crops_follow_reference_center_crop():
    """A.Scale((128,171)) + A.CenterCrop(112) + ToTensor of the evaluation / retrieval transforms (classifier.py:683-695,
    809-817), written with PIL exactly as utils/augmentation.py:185-191 does, equals oracle.scale_crop at center_crops()."""
    from PIL import Image
    from torchvision import transforms
    from dualvar_b200 import frames as FR
    from oracle.frames import scale_crop
    cc = FR.center_crops(2, 1)
    assert cc.shape == (2, 1, 2) and cc[0, 0].tolist() == [8, 30]            # round(29.5) == 30: half to even
    assert FR.center_crops(1, 1, scaled=(128, 171), crop=(113, 112))[0, 0].tolist() == [8, 30]   # round(7.5) == 8
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (2, 3, 90, 120, 3), dtype=np.uint8)
    got = scale_crop(frames, cc.numpy(), 1)
    for b in range(2):
        for f in range(3):
            im = Image.fromarray(frames[b, f]).resize((128, 171), Image.BICUBIC)
            w, h = im.size
            th, tw = 112, 112
            x1, y1 = int(round((w - tw) / 2.)), int(round((h - th) / 2.))
            ten = transforms.ToTensor()(im.crop((x1, y1, x1 + tw, y1 + th)))
            assert torch.equal(ten, torch.from_numpy(got[b, :, f]).float() / 255)
