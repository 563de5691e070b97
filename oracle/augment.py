"""Oracle for the colour-jitter stage of the reference loader (second stage): ``A.ColorJitter`` of
utils/augmentation.py:429-660 as pretrain.py:505 builds it (block = 1): for every frame - or once per clip when
``consistent`` - with probability p a fresh random transform made of torchvision's tensor ``adjust_brightness /
adjust_contrast / adjust_saturation / adjust_hue`` in a shuffled order, applied to the ToTensor output (CHW float32 in [0,1]).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg): the product path never imports
this module.

The arithmetic lives in torchvision (third-party, 0.26.0 in this image; torchvision/transforms/_functional_tensor.py:
_blend, rgb_to_grayscale, _rgb2hsv, _hsv2rgb), restated here operation by operation in float32 so that the CUDA kernel can
follow the same rounding sequence. Pinned against torchvision itself and against the real reference class
(tests/golden/color_jitter.npz, written by tests/golden/make_golden_color_jitter.py) in tests/test_color_jitter.py.
"""
import math

import numpy as np
import torch

OPS = ("brightness", "contrast", "saturation", "hue")


def _f32(x):
    return torch.tensor(float(x), dtype=torch.float32)


def _gray(img):
    r, g, b = img[0], img[1], img[2]
    return (_f32(0.2989) * r + _f32(0.587) * g) + _f32(0.114) * b


def _blend(a, b, ratio, one_minus=None):
    """torchvision _blend: (ratio * a + (1 - ratio) * b).clamp(0, 1). Both scalars reach the float32 kernels as float32;
    1 - ratio is formed in double first (``one_minus`` carries that value when ``ratio`` is already rounded)."""
    om = 1.0 - float(ratio) if one_minus is None else one_minus
    return (_f32(ratio) * a + _f32(om) * b).clamp(0.0, 1.0)


def adjust_brightness(img, f, om=None):
    return _blend(img, torch.zeros_like(img), f, om)


def adjust_contrast(img, f, om=None):
    mean = _gray(img).mean()
    return _blend(img, mean, f, om)


def adjust_saturation(img, f, om=None):
    return _blend(img, _gray(img).unsqueeze(0), f, om)


def adjust_hue(img, f, om=None):
    r, g, b = img[0], img[1], img[2]
    maxc = torch.maximum(torch.maximum(r, g), b)
    minc = torch.minimum(torch.minimum(r, g), b)
    eqc = maxc == minc
    cr = maxc - minc
    ones = torch.ones_like(maxc)
    s = cr / torch.where(eqc, ones, maxc)
    div = torch.where(eqc, ones, cr)
    rc, gc, bc = (maxc - r) / div, (maxc - g) / div, (maxc - b) / div
    hr = (maxc == r).float() * (bc - gc)
    hg = ((maxc == g) & (maxc != r)).float() * ((2.0 + rc) - bc)
    hb = ((maxc != g) & (maxc != r)).float() * ((4.0 + gc) - rc)
    h = (hr + hg) + hb
    h = torch.fmod(h / 6.0 + 1.0, 1.0)
    h = torch.remainder(h + _f32(f), 1.0)
    v = maxc
    h6 = h * 6.0
    i = torch.floor(h6)
    fr = h6 - i
    i = i.to(torch.int32) % 6
    p = (v * (1.0 - s)).clamp(0.0, 1.0)
    q = (v * (1.0 - s * fr)).clamp(0.0, 1.0)
    t = (v * (1.0 - s * (1.0 - fr))).clamp(0.0, 1.0)
    sel = lambda opts: sum((i == k).float() * o for k, o in enumerate(opts))  # noqa: E731
    return torch.stack((sel((v, q, p, p, t, v)), sel((t, v, v, q, p, p)), sel((p, p, t, v, v, q))))


_FN = {0: adjust_brightness, 1: adjust_contrast, 2: adjust_saturation, 3: adjust_hue}


def draw_color_jitter(n_frames, py_random, np_random, brightness=0.8, contrast=0.8, saturation=0.8, hue=0.2, p=0.8,
                      consistent=False, seq_len=16):
    """The random draws of A.ColorJitter.__call__ (block = 1) for a list of n_frames images, in its order:
    per frame (per seq_len frames when consistent) ``np.random.uniform(0, 1) < p`` decides whether a transform is
    drawn; get_params then takes ``random.uniform`` for brightness, contrast, saturation, hue and ``random.shuffle``s the
    four ops (utils/augmentation.py:480-510,595-599). Returns float32 [n_frames][12]:
    apply, b, 1-b, c, 1-c, s, 1-s, h, op0..op3 (op codes 0..3 = brightness, contrast, saturation, hue)."""
    lo = lambda v: max(0.0, 1.0 - v)  # noqa: E731
    out = np.zeros((n_frames, 12), np.float32)
    cur = None
    for idx in range(n_frames):
        if not consistent or idx % seq_len == 0:
            if np_random.uniform(0., 1.) < p:
                fb = py_random.uniform(lo(brightness), 1.0 + brightness)
                fc = py_random.uniform(lo(contrast), 1.0 + contrast)
                fs = py_random.uniform(lo(saturation), 1.0 + saturation)
                fh = py_random.uniform(-hue, hue)
                order = [0, 1, 2, 3]
                py_random.shuffle(order)
                cur = np.array([1.0, fb, 1.0 - fb, fc, 1.0 - fc, fs, 1.0 - fs, fh] + order, np.float64)
            else:
                cur = np.zeros(12, np.float64)
        out[idx] = cur.astype(np.float32)
    return out


def color_jitter(frames, params):
    """frames: (n_frames, 3, H, W) float32 in [0, 1] (ToTensor output); params from draw_color_jitter."""
    out = []
    for x, prm in zip(frames, params):
        if prm[0] != 0:
            fac = {0: (prm[1], prm[2]), 1: (prm[3], prm[4]), 2: (prm[5], prm[6]), 3: (prm[7], None)}
            for op in prm[8:12].astype(int):
                f, om = fac[int(op)]
                x = _FN[int(op)](x, float(f), None if om is None else float(om))
        out.append(x)
    return torch.stack(out)


# ----------------------------------------------------------------------------------------------- Gaussian blur
# A.GaussianBlur (utils/augmentation.py:706-721): per clip sigma = random.uniform(0.1, 2); every frame goes
# ToPILImage -> PIL ImageFilter.GaussianBlur(radius=sigma) -> ToTensor. Pillow (12.2.0; src/libImaging/BoxBlur.c)
# approximates the Gaussian with three passes of an "extended box" filter per direction; restated from the published
# algorithm and pinned bit for bit against Pillow itself in tests/test_stage_blur.py.
_f = np.float32


def gaussian_box_radius(radius, passes=3):
    """BoxBlur.c _gaussian_blur_radius (float variables, double literals: the C expression types are kept)."""
    radius = _f(radius)
    sigma2 = _f(_f(radius * radius) / _f(passes))
    L = _f(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = _f(math.floor((float(L) - 1.0) / 2.0))                                              # noqa: E741
    a = _f(float(_f(_f(2) * l + _f(1))) * (float(_f(l * _f(l + _f(1)))) - 3.0 * float(sigma2)))
    a = _f(float(a) / (6.0 * float(_f(sigma2 - _f(_f(l + _f(1)) * _f(l + _f(1)))))))
    return _f(l + a)


def _line_box_blur(line, radius, ww, fw):
    """ImagingLineBoxBlur32 on an (n, C) int64 array of uint8 values; uint32 wrap-around arithmetic."""
    n = line.shape[0]
    lastx = n - 1
    edge_a, edge_b = min(radius + 1, n), max(n - radius - 1, 0)
    M = 1 << 32
    out = np.zeros_like(line)
    acc = (line[0] * (radius + 1)) % M
    for x in range(edge_a - 1):
        acc = (acc + line[x]) % M
    acc = (acc + line[lastx] * (radius - edge_a + 1)) % M

    def step(x, sub, add, left, right):
        nonlocal acc
        acc = (acc + line[add] - line[sub]) % M
        bulk = (acc * ww + (line[left] + line[right]) * fw) % M
        out[x] = ((bulk + (1 << 23)) % M) >> 24

    if edge_a <= edge_b:
        for x in range(0, edge_a):
            step(x, 0, x + radius, 0, x + radius + 1)
        for x in range(edge_a, edge_b):
            step(x, x - radius - 1, x + radius, x - radius - 1, x + radius + 1)
        for x in range(edge_b, lastx + 1):
            step(x, x - radius - 1, lastx, x - radius - 1, lastx)
    else:
        for x in range(0, edge_b):
            step(x, 0, x + radius, 0, x + radius + 1)
        for x in range(edge_b, edge_a):
            step(x, 0, lastx, 0, lastx)
        for x in range(edge_a, lastx + 1):
            step(x, x - radius - 1, lastx, x - radius - 1, lastx)
    return out


def _horizontal_box_blur(img, float_radius):
    radius = int(float_radius)
    ww = int(_f(_f(1 << 24) / _f(float_radius * _f(2) + _f(1))))
    fw = (((1 << 24) - (radius * 2 + 1) * ww) % (1 << 32)) // 2
    src = img.astype(np.int64)
    out = np.empty_like(src)
    for y in range(img.shape[0]):
        out[y] = _line_box_blur(src[y], radius, ww, fw)
    return out.astype(np.uint8)


def pil_gaussian_blur(img, sigma, passes=3):
    """``Image.fromarray(img).filter(ImageFilter.GaussianBlur(radius=sigma))`` for an (H, W, 3) uint8 array."""
    if sigma == 0:
        return img.copy()
    r = gaussian_box_radius(sigma, passes)
    out = img
    if r != 0:
        for _ in range(passes):
            out = _horizontal_box_blur(out, r)
        t = out.transpose(1, 0, 2)
        for _ in range(passes):
            t = _horizontal_box_blur(t, r)
        out = t.transpose(1, 0, 2)
    return out


def draw_gaussian_blur(n_frames, py_random, sigma=(0.1, 2.0), seq_len=16):
    """Per-frame sigma in A.GaussianBlur's draw order: one ``random.uniform(sigma[0], sigma[1])`` per seq_len frames."""
    out, cur = [], 0.0
    for idx in range(n_frames):
        if idx % seq_len == 0:
            cur = py_random.uniform(sigma[0], sigma[1])
        out.append(cur)
    return out


def gaussian_blur(frames, sigmas):
    """frames: (n, 3, H, W) float32 in [0, 1]; sigma <= 0 leaves a frame untouched (stage skipped)."""
    out = []
    for x, sg in zip(frames, sigmas):
        if sg > 0:
            u8 = x.mul(255).byte().permute(1, 2, 0).numpy()                      # ToPILImage
            x = torch.from_numpy(pil_gaussian_blur(u8, sg)).permute(2, 0, 1).float().div(255)      # ToTensor
        out.append(x)
    return torch.stack(out)
