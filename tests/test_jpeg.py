"""JPEG frame decoding (reference: PIL.Image.open per frame, dataset/local_dataset.py:283-286).

CPU: the oracle (oracle/jpeg.py, a numpy restatement of libjpeg's islow IDCT / fancy upsampling / colour conversion) is
pinned bit for bit against Pillow itself and against a committed fixture; the product's host half (Huffman decoding in the
C-ABI library, no GPU needed) is checked against the oracle coefficient for coefficient.
GPU: dualvar_b200.jpeg.decode_batch == Pillow, pixel for pixel; decode -> Scale -> crop chain == the PIL chain."""
import io
import os

import numpy as np
import pytest
import torch

from oracle import jpeg as OJ

PIL = pytest.importorskip("PIL.Image")


def _smooth(rng, h, w):
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3)).astype(np.uint8)
    im = np.asarray(PIL.fromarray(base).resize((w, h), PIL.BICUBIC)).astype(np.float32)
    return np.clip(im + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)


def _encode(img, **kw):
    b = io.BytesIO()
    PIL.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


def _pil(data):
    return np.asarray(PIL.open(io.BytesIO(data)).convert("RGB"))


CASES = [(h, w, sub, q) for (h, w) in [(16, 16), (24, 40), (37, 53), (8, 8), (17, 9), (1, 1), (65, 130)]
         for sub in (0, 1, 2) for q in (30, 90)] + [(240, 320, 2, 75)]


def test_oracle_is_bit_exact_with_pillow():
    rng = np.random.default_rng(0)
    for h, w, sub, q in CASES:
        data = _encode(_smooth(rng, h, w), quality=q, subsampling=sub)
        assert np.array_equal(OJ.decode(data), _pil(data)), (h, w, sub, q)
    img = _smooth(rng, 48, 64)
    d = _encode(img, quality=80, subsampling=2, restart_marker_blocks=3)
    assert b"\xff\xdd" in d and np.array_equal(OJ.decode(d), _pil(d))                     # restart intervals
    d = _encode(img[:, :, 0], quality=80)
    assert np.array_equal(OJ.decode(d), _pil(d))                                           # grayscale
    with pytest.raises(ValueError):
        OJ.decode(_encode(img, quality=80, progressive=True))                              # not baseline: refused


def test_oracle_matches_committed_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "jpeg.npz"))
    for i in range(int(g["n"])):
        assert np.array_equal(OJ.decode(g[f"file{i}"].tobytes()), g[f"rgb{i}"]), i


def test_host_huffman_decoder_matches_oracle():
    """The C-ABI library's host half: headers, Huffman decoding, de-zigzag - every coefficient and table."""
    from dualvar_b200 import _lib, jpeg as JP
    rng = np.random.default_rng(1)
    for h, w, sub, q in [(48, 64, 2, 50), (48, 64, 1, 90), (48, 64, 0, 75), (37, 53, 2, 95), (240, 320, 2, 75)]:
        files = [_encode(_smooth(rng, h, w), quality=q, subsampling=sub) for _ in range(3)]
        coef, qt, info = JP.decode_host_coefficients(files, threads=2)
        for i, f in enumerate(files):
            oc, hd = OJ.decode_coefficients(f)
            assert (info["width"], info["height"], info["hmax"], info["vmax"]) == (w, h, hd["hmax"], hd["vmax"])
            assert np.array_equal(np.concatenate([c.reshape(-1) for c in oc]), coef[i].numpy())
            for ci, c in enumerate(hd["comps"]):
                assert np.array_equal(hd["qt"][c[3]], qt[i, ci].numpy().astype(np.int32))
    f = _encode(_smooth(rng, 48, 64), quality=80, subsampling=2, restart_marker_blocks=2)
    coef, _, _ = JP.decode_host_coefficients([f])
    assert np.array_equal(np.concatenate([c.reshape(-1) for c in OJ.decode_coefficients(f)[0]]), coef[0].numpy())
    gray = _encode(_smooth(rng, 40, 40)[:, :, 0], quality=70)
    coef, _, info = JP.decode_host_coefficients([gray])
    assert info["components"] == 1 and np.array_equal(OJ.decode_coefficients(gray)[0][0].reshape(-1), coef[0].numpy())
    # refused inputs come back as status codes with a message, nothing crashes
    with pytest.raises(_lib.DualVarNativeError, match="baseline"):
        JP.probe(_encode(_smooth(rng, 32, 32), quality=80, progressive=True))
    with pytest.raises(_lib.DualVarNativeError, match="SOI"):
        JP.probe(b"not a jpeg at all")
    with pytest.raises(_lib.DualVarNativeError, match="share size"):
        JP.decode_host_coefficients([files[0], _encode(_smooth(rng, 32, 32), quality=80)])
    with pytest.raises(_lib.DualVarNativeError):
        JP.decode_host_coefficients([files[0][:len(files[0]) // 2][:200] + b"\\xff\\xd9"])


def test_decode_refuses_cpu_device():
    from dualvar_b200 import _lib, jpeg as JP
    with pytest.raises(_lib.DualVarNativeError):
        JP.decode_batch([_encode(np.zeros((8, 8, 3), np.uint8))], "cpu")


@pytest.mark.gpu
def test_gpu_decode_is_bit_exact_with_pillow():
    from dualvar_b200 import jpeg as JP
    rng = np.random.default_rng(2)
    for h, w, sub, q in CASES + [(112, 112, 2, 60)]:
        files = [_encode(_smooth(rng, h, w), quality=q, subsampling=sub) for _ in range(3)]
        got = JP.decode_batch(files, "cuda:0").cpu().numpy()
        for i, f in enumerate(files):
            assert np.array_equal(got[i], _pil(f)), (h, w, sub, q, i)
    gray = [_encode(_smooth(rng, 40, 56)[:, :, 0], quality=70)]
    assert np.array_equal(JP.decode_batch(gray, "cuda:0").cpu().numpy()[0], _pil(gray[0]))
    rst = [_encode(_smooth(rng, 48, 64), quality=80, subsampling=2, restart_marker_blocks=2)]
    assert np.array_equal(JP.decode_batch(rst, "cuda:0").cpu().numpy()[0], _pil(rst[0]))


@pytest.mark.gpu
def test_gpu_decode_scale_crop_chain_matches_pil_chain():
    """JPEG bytes -> decode -> Scale((128,171)) bicubic -> crop on the GPU == Image.open -> resize -> crop with Pillow."""
    from dualvar_b200 import frames as FR, jpeg as JP
    rng = np.random.default_rng(3)
    B, F = 2, 4
    files = [_encode(_smooth(rng, 120, 160), quality=85, subsampling=2) for _ in range(B * F)]
    fr = JP.decode_batch(files, "cuda:0").view(B, F, 120, 160, 3)
    crops = torch.tensor([[[5, 20]], [[16, 59]]], dtype=torch.int32)
    got = FR.scale_crop(fr, crops, 1).cpu().numpy()                     # (B, 3, F, 112, 112)
    for b in range(B):
        for f in range(F):
            im = PIL.open(io.BytesIO(files[b * F + f])).resize((128, 171), PIL.BICUBIC)
            l, t = int(crops[b, 0, 0]), int(crops[b, 0, 1])
            want = np.asarray(im.crop((l, t, l + 112, t + 112))).transpose(2, 0, 1)
            assert np.array_equal(got[b, :, f], want), (b, f)
