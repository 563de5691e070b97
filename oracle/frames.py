"""Oracle for the frame-staging row next to the hot path (first stage): the first
stages of the reference's augmentation chain - ``A.Scale((128, 171))`` (PIL bicubic), ``A.RandomCrop(112)``, ``A.ToTensor()``
(utils/augmentation.py:125-176,361-364; the ``null_transform`` of pretrain.py:491-497) - restated in numpy.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg): the product path never imports
this module.

The arithmetic lives in Pillow (third-party, not under /root/reference; 12.2.0 in this image), whose 8-bit resampler
(src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
Vertical_8bpc) is restated here from its published algorithm: per output pixel a window of the bicubic kernel
(a = -0.5, support 2 x max(scale, 1)) evaluated in double, normalised, converted to 22-bit fixed point with
round-half-away-from-zero, accumulated in int32 from 1 << 21, shifted and saturated; horizontal pass first, through a
uint8 intermediate, then vertical. Pinned bit-for-bit against Pillow itself in tests/test_frames.py.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size, out_size):
    """(xmin[out], xmax[out], kk[out][ksize] int32) of one axis - Resample.c precompute_coeffs + normalize_coeffs_8bpc."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, np.int32)
    xmax_a = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            v = w * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w < 0 else int(0.5 + v)
        xmin_a[xx], xmax_a[xx] = xmin, xmax
    return xmin_a, xmax_a, kk


def _pass(img, out_size, axis):
    """One 8-bit resampling pass along ``axis`` (0 = vertical, 1 = horizontal) of an (H, W, C) uint8 image."""
    xmin, xmax, kk = resample_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(xmax[xx]):
            acc += src[xmin[xx] + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_bicubic_resize(img, out_w, out_h):
    """``Image.fromarray(img).resize((out_w, out_h), Image.BICUBIC)`` for an (H, W, 3) uint8 array."""
    h, w = img.shape[:2]
    if w != out_w:
        img = _pass(img, out_w, 1)
    if h != out_h:
        img = _pass(img, out_h, 0)
    return img


def scale_crop(frames, crops, n_views, scale_size=(128, 171), crop_size=(112, 112)):
    """frames: (B, F, Hs, Ws, 3) uint8 decoded frames, F = n_views * T; crops: (B, n_views, 2) int (h_start, w_start) as
    RandomCrop draws them. Returns (B, 3, F, 112, 112) uint8: Scale(scale_size) then
    ``img.crop((h_start, w_start, h_start + size[0], w_start + size[1]))`` per frame (utils/augmentation.py:131-176; the
    reference hands its (128, 171) to PIL as (width, height) and its "h" offsets to crop() as the left edge).
    ToTensor is the uint8 -> x / 255 conversion the ingest applies (oracle: torch / numpy division)."""
    B, F = frames.shape[:2]
    T = F // n_views
    out = np.empty((B, 3, F, crop_size[1], crop_size[0]), np.uint8)
    for b in range(B):
        for f in range(F):
            r = pil_bicubic_resize(frames[b, f], scale_size[0], scale_size[1])
            left, upper = int(crops[b, f // T, 0]), int(crops[b, f // T, 1])
            out[b, :, f] = r[upper:upper + crop_size[1], left:left + crop_size[0]].transpose(2, 0, 1)
    return out


def draw_crops(B, n_views, rng, scaled=(128, 171), crop=(112, 112)):
    """The (h_start, w_start) pairs in the order A.RandomCrop draws them with ``random.randint`` (one pair per clip,
    utils/augmentation.py:164-166: h = img.size[0] = width, w = img.size[1] = height)."""
    out = np.zeros((B, n_views, 2), np.int32)
    for b in range(B):
        for v in range(n_views):
            out[b, v, 0] = rng.randint(0, scaled[0] - crop[0])
            out[b, v, 1] = rng.randint(0, scaled[1] - crop[1])
    return out
