"""Frame staging in front of the encoders (first stage): the first stages of the reference loader's
per-clip transform - ``A.Scale((128, 171))`` + ``A.RandomCrop(112)`` + ``A.ToTensor()`` (the ``null_transform`` of
pretrain.py:491-497, utils/augmentation.py:125-176,361-364) - on the GPU from decoded uint8 frames, bit-exact with
Pillow's bicubic resampler. The reference does this on the host CPU with PIL for every one of the 48 frames of a sample
(dataset/local_dataset.py:289-300), which at ~850 samples/s per GPU is the first thing to starve the step.

    crops = frames.draw_crops(B, 3)                                   # random.randint in the reference's order
    clips = frames.stage_clips(frames_u8.cuda(), crops, n_views=3)    # engine.RawClips over uint8 (B,3,48,112,112)
    ret = model(clips)

Second stage: ``A.ColorJitter`` (utils/augmentation.py:429-660, the jitter of ``base_transform`` /
``same_series_transform``, pretrain.py:505) - ``draw_color_jitter`` draws the per-frame factors and op order in the
reference's RNG order, ``color_jitter`` / ``stage_clips(..., jitter=...)`` apply them on the GPU (one CTA per frame, frame
resident in shared memory). Third stage: ``A.GaussianBlur`` (:706-721: ToPILImage -> PIL's Gaussian -> ToTensor) -
``draw_gaussian_blur`` / ``gaussian_blur`` / ``stage_clips(..., blur=...)``; its line filter is one host/device function
that the CPU tests pin against Pillow through ``dv_frames_gaussian_blur_host``.
"""
import ctypes
import random

import torch

from . import _lib
from .engine import RawClips

call, ptr, stream_ptr = _lib.call, _lib.ptr, _lib.stream_ptr


def draw_crops(B, n_views, rng=random, scaled=(128, 171), crop=(112, 112)):
    """(B, n_views, 2) int32 (h_start, w_start) pairs in the order A.RandomCrop draws them: one pair per clip,
    ``random.randint(0, h - size0)`` then ``random.randint(0, w - size1)`` with h = img.size[0] = 128 (the PIL width) and
    w = img.size[1] = 171 (utils/augmentation.py:160-166). crop() takes them as (left, upper)."""
    hi0, hi1 = scaled[0] - crop[0], scaled[1] - crop[1]
    vals = [(rng.randint(0, hi0), rng.randint(0, hi1)) for _ in range(B * n_views)]
    return torch.tensor(vals, dtype=torch.int32).view(B, n_views, 2)


def center_crops(B, n_views, scaled=(128, 171), crop=(112, 112)):
    """(B, n_views, 2) int32 (left, upper) of ``A.CenterCrop`` (utils/augmentation.py:185-191: ``int(round((w - tw) / 2.))``,
    Python's round-half-to-even) - the crop of the evaluation / retrieval transforms (classifier.py:683-695,809-817)."""
    left, upper = int(round((scaled[0] - crop[0]) / 2.)), int(round((scaled[1] - crop[1]) / 2.))
    return torch.tensor([left, upper], dtype=torch.int32).repeat(B, n_views, 1)


def scale_crop(frames, crops, n_views, scale_size=(128, 171), crop_size=(112, 112)):
    """frames: uint8 CUDA tensor (B, n_views*T, Hs, Ws, 3) of decoded frames; crops: int32 (B, n_views, 2).
    Returns uint8 (B, 3, n_views*T, crop_h, crop_w): every frame resized to scale_size = (width, height) with PIL's
    bicubic filter and cropped at its clip's (left, upper)."""
    if not frames.is_cuda:
        raise _lib.DualVarNativeError("scale_crop: frames must be on a B200 (no CPU fallback)")
    assert frames.dtype == torch.uint8 and frames.dim() == 5 and frames.shape[4] == 3, "frames: uint8 (B, F, H, W, 3)"
    frames = frames.contiguous()
    B, F, Hs, Ws, _ = frames.shape
    assert F % n_views == 0
    crops = crops.to(device=frames.device, dtype=torch.int32).contiguous()
    assert tuple(crops.shape) == (B, n_views, 2)
    sw, sh = scale_size
    cw, ch = crop_size
    tmp = torch.empty((B * F, Hs, sw, 4), dtype=torch.uint8, device=frames.device)      # RGBX intermediate
    out = torch.empty((B, 3, F, ch, cw), dtype=torch.uint8, device=frames.device)
    call("dv_frames_scale_crop_u8", ptr(frames), ptr(tmp), ptr(out), ptr(crops), B, n_views, F // n_views, Hs, Ws, sw, sh,
         cw, ch, stream_ptr())
    return out


def _jitter_rows(n_frames, py_random, np_random, brightness, contrast, saturation, hue, p, consistent, seq_len):
    lo = lambda v: max(0.0, 1.0 - v)  # noqa: E731
    b0, b1, c0, c1, s0, s1 = lo(brightness), 1.0 + brightness, lo(contrast), 1.0 + contrast, lo(saturation), 1.0 + saturation
    zero = [0.0] * 12
    rows, cur = [], None
    for idx in range(n_frames):
        if not consistent or idx % seq_len == 0:
            if np_random.uniform(0., 1.) < p:
                fb = py_random.uniform(b0, b1)
                fc = py_random.uniform(c0, c1)
                fs = py_random.uniform(s0, s1)
                fh = py_random.uniform(-hue, hue)
                order = [0.0, 1.0, 2.0, 3.0]
                py_random.shuffle(order)
                cur = [1.0, fb, 1.0 - fb, fc, 1.0 - fc, fs, 1.0 - fs, fh] + order
            else:
                cur = zero
        rows.append(cur)
    return rows


def draw_color_jitter(n_frames, py_random=random, np_random=None, brightness=0.8, contrast=0.8, saturation=0.8, hue=0.2,
                      p=0.8, consistent=False, seq_len=16):
    """The random draws of ``A.ColorJitter.__call__`` (block = 1) for a list of n_frames images, in its order: per frame
    (per seq_len frames when ``consistent``) ``np.random.uniform(0, 1) < p`` decides whether a transform is drawn;
    ``get_params`` then takes ``random.uniform`` for brightness, contrast, saturation, hue and ``random.shuffle``s the four
    ops (utils/augmentation.py:480-510,595-599). Returns float32 (n_frames, 12): apply, b, 1-b, c, 1-c, s, 1-s, h,
    op0..op3 with op codes 0..3 = brightness, contrast, saturation, hue (1 - x formed in double, as torchvision's _blend does)."""
    import numpy as np
    np_random = np.random if np_random is None else np_random
    rows = _jitter_rows(n_frames, py_random, np_random, brightness, contrast, saturation, hue, p, consistent, seq_len)
    return torch.tensor(rows, dtype=torch.float64).to(torch.float32)


def color_jitter(clips_u8, params):
    """clips_u8: uint8 CUDA tensor (B, 3, F, H, W) (scale_crop's output); params: float32 (B*F, 12) from draw_color_jitter
    (frame order b-major, as the loader walks a sample's frames). Returns float32 (B, 3, F, H, W) in [0, 1]:
    ToTensor followed by the frame's jitter - what the reference hands to Normalize."""
    if not clips_u8.is_cuda:
        raise _lib.DualVarNativeError("color_jitter: clips must be on a B200 (no CPU fallback)")
    assert clips_u8.dtype == torch.uint8 and clips_u8.dim() == 5 and clips_u8.shape[1] == 3
    clips_u8 = clips_u8.contiguous()
    B, _, F, H, W = clips_u8.shape
    params = params.to(device=clips_u8.device, dtype=torch.float32).contiguous()
    assert tuple(params.shape) == (B * F, 12)
    out = torch.empty((B, 3, F, H, W), dtype=torch.float32, device=clips_u8.device)
    call("dv_frames_color_jitter", ptr(clips_u8), ptr(out), ptr(params), B, F, H, W, stream_ptr())
    return out


def draw_gaussian_blur(n_frames, py_random=random, sigma=(0.1, 2.0), seq_len=16):
    """Per-frame sigma in ``A.GaussianBlur``'s draw order (utils/augmentation.py:713-718): one
    ``random.uniform(sigma[0], sigma[1])`` per seq_len frames. Set the entries of clips whose RandomApply skipped the
    stage to 0."""
    out, cur = [], 0.0
    for idx in range(n_frames):
        if idx % seq_len == 0:
            cur = py_random.uniform(sigma[0], sigma[1])
        out.append(cur)
    return out


def blur_params(sigmas):
    """int32 (n_frames, 4) = {apply, box radius, ww, fw}: the fixed-point extended-box parameters Pillow derives from each
    sigma (computed by the library's host code, the same the CPU test pins against Pillow); sigma <= 0 -> not applied."""
    buf = (ctypes.c_int32 * 4)()
    seen, rows = {}, []
    for sg in sigmas:                       # one sigma per clip: a few hundred distinct values per batch of frames
        sg = float(sg)
        row = seen.get(sg)
        if row is None:
            call("dv_frames_gaussian_blur_params_host", ctypes.c_float(sg), buf)
            row = seen[sg] = list(buf)
        rows.append(row)
    return torch.tensor(rows, dtype=torch.int32)


def gaussian_blur(clips, sigmas):
    """clips: float32 CUDA tensor (B, 3, F, H, W) in [0, 1] (ToTensor / colour-jitter output); sigmas: B*F values (0 = frame
    not blurred). Returns a new float32 tensor: every blurred frame went ToPILImage -> PIL GaussianBlur -> ToTensor."""
    if not clips.is_cuda:
        raise _lib.DualVarNativeError("gaussian_blur: clips must be on a B200 (no CPU fallback)")
    assert clips.dtype == torch.float32 and clips.dim() == 5 and clips.shape[1] == 3
    clips = clips.contiguous()
    B, _, F, H, W = clips.shape
    prm = blur_params(sigmas).to(clips.device)
    assert tuple(prm.shape) == (B * F, 4)
    out = torch.empty_like(clips)
    call("dv_frames_gaussian_blur", ptr(clips), ptr(out), ptr(prm), B, F, H, W, stream_ptr())
    return out


def gaussian_blur_host(frame, sigma):
    """The same stage on one float32 CHW frame on the host, running the line filter the kernel runs (CPU parity tests)."""
    frame = frame.contiguous().float()
    assert not frame.is_cuda and frame.dim() == 3 and frame.shape[0] == 3
    out = torch.empty_like(frame)
    call("dv_frames_gaussian_blur_host", ctypes.c_void_p(frame.data_ptr()), ctypes.c_void_p(out.data_ptr()),
         frame.shape[1], frame.shape[2], ctypes.c_float(float(sigma)))
    return out


def draw_plan(n_samples, n_views=3, seq_len=16, scaled=(128, 171), crop=(112, 112),
              weights=((0.2, 0.8, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)), jitter_p=0.8, blur_p=0.5, consistent=False,
              py_random=random, np_random=None):
    """Every random draw of the loader's transform for n_samples samples, in the reference's order and from the RNGs it
    uses (pretrain.py:491-529): per clip ``MultiRandomizedTransform`` picks a branch with ``np.random.uniform()``
    (utils/augmentation.py:795-809; branch 0 = null_transform, 1 = base, 2 = same-series), ``RandomCrop`` takes two
    ``random.randint``; in the jitter branches torchvision's ``RandomApply`` consumes ``torch.rand(1)`` (global generator)
    before ``ColorJitter`` (p = 0.8) and again before ``GaussianBlur`` (p = 0.5), each of which then draws as
    draw_color_jitter / draw_gaussian_blur describe. Returns {'crops': int32 (n, V, 2), 'jitter': float32 (n*V*T, 12),
    'blur': [n*V*T sigmas], 'branch': int32 (n, V)} - the arguments of ``stage_clips``; frames of clips whose stage was
    skipped carry apply = 0 / sigma = 0. Pure host work (~0.4 ms per sample, the cost of Python's own RNG calls): it belongs
    where the reference runs its transform, in the DataLoader workers, next to the JPEG decode."""
    import numpy as np
    np_random = np.random if np_random is None else np_random
    cum = [np.cumsum(w) for w in weights]
    crops, branches, jitter, blur = [], [], [], []
    zero_rows = [[0.0] * 12] * seq_len
    for _ in range(n_samples):
        for v in range(n_views):
            rand_p = np_random.uniform()
            ind = 0
            while rand_p >= cum[v][ind]:
                ind += 1
            branches.append(ind)
            crops.append((py_random.randint(0, scaled[0] - crop[0]), py_random.randint(0, scaled[1] - crop[1])))
            jit = zero_rows
            sig = [0.0] * seq_len
            if ind != 0:
                if not (jitter_p < float(torch.rand(1))):            # RandomApply([ColorJitter], p=0.8)
                    jit = _jitter_rows(seq_len, py_random, np_random, 0.8, 0.8, 0.8, 0.2, 0.8, consistent, seq_len)
                if not (blur_p < float(torch.rand(1))):              # RandomApply([GaussianBlur], p=0.5)
                    sig = draw_gaussian_blur(seq_len, py_random, seq_len=seq_len)
            jitter.extend(jit)
            blur.extend(sig)
    return {"crops": torch.tensor(crops, dtype=torch.int32).view(n_samples, n_views, 2),
            "jitter": torch.tensor(jitter, dtype=torch.float64).to(torch.float32), "blur": blur,
            "branch": torch.tensor(branches, dtype=torch.int32).view(n_samples, n_views)}


def stage_clips(frames, crops, n_views, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), jitter=None, blur=None):
    """Decoded frames -> the model input: Scale + RandomCrop (+ ColorJitter when ``jitter`` = draw_color_jitter(...)
    parameters are given, + GaussianBlur when ``blur`` = per-frame sigmas from draw_gaussian_blur(...)) here,
    ToTensor + Normalize + layout in the ingest kernel."""
    clips = scale_crop(frames, crops, n_views)
    if jitter is not None or blur is not None:
        if jitter is None:        # ToTensor only: a parameter table with every `apply` flag clear
            jitter = torch.zeros((clips.shape[0] * clips.shape[2], 12), dtype=torch.float32)
        clips = color_jitter(clips, jitter)
    if blur is not None:
        clips = gaussian_blur(clips, blur)
    return RawClips(clips, n_views, mean, std)


def axis_table(in_size, out_size):
    """Host-side fixed-point bicubic table of one axis as the library computes it: (ksize, int32 [out_size][ksize+2])."""
    cap = out_size * (4 * max(1, -(-in_size // out_size)) + 8)
    buf = (ctypes.c_int32 * cap)()
    ks = ctypes.c_int32(0)
    call("dv_frames_axis_table_host", in_size, out_size, buf, cap, ctypes.byref(ks))
    t = torch.tensor(list(buf[:out_size * (ks.value + 2)]), dtype=torch.int32).view(out_size, ks.value + 2)
    return ks.value, t
