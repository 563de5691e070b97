"""JPEG frames -> decoded uint8 frames on the GPU (reference: PIL.Image.open per frame in the loader,
dataset/local_dataset.py:283-286, i.e. libjpeg-turbo on DataLoader workers).

``decode_batch(files, device)`` takes the raw bytes of n baseline JPEG files of one geometry (the frames of a clip batch)
and returns uint8 (n, H, W, 3) on ``device``, bit-identical to ``numpy.asarray(PIL.Image.open(f).convert('RGB'))``:
entropy decoding runs on host threads inside the C-ABI library (the bit stream is serial), the quantised coefficients go
to the device through pinned memory (2 bytes per sample - no more than the decoded pixels would take) and dequantisation,
inverse DCT, chroma upsampling and colour conversion run as two kernels over the whole batch (csrc/jpeg.cu).
``decode_host_coefficients`` exposes the host half alone (CPU tests pin it against the oracle without a GPU).
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ptr, stream_ptr

_pinned = {}


def probe(data):
    """{'width','height','components','hmax','vmax','mcux','mcuy','coef_count'} of one JPEG file (bytes)."""
    info = (ctypes.c_int32 * 8)()
    cc = ctypes.c_int64(0)
    buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
    _lib.check(_lib.load().dv_jpeg_probe_host(buf, len(data), info, ctypes.byref(cc)), "dv_jpeg_probe_host")
    keys = ("width", "height", "components", "hmax", "vmax", "mcux", "mcuy")
    out = {k: int(info[i]) for i, k in enumerate(keys)}
    out["coef_count"] = int(cc.value)
    out["_info"] = info
    return out


def _huffman(files, info, coef, qt, threads):
    n = len(files)
    bufs = [(ctypes.c_uint8 * len(f)).from_buffer_copy(f) for f in files]
    ptrs = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
    lens = (ctypes.c_int64 * n)(*[len(f) for f in files])
    _lib.check(_lib.load().dv_jpeg_huffman_decode_host(ptrs, lens, n, info["_info"], ctypes.c_void_p(coef.data_ptr()),
                                                       info["coef_count"], ctypes.c_void_p(qt.data_ptr()), threads),
               "dv_jpeg_huffman_decode_host")


def decode_host_coefficients(files, threads=None):
    """Host half only: (int16 [n][coef_count] coefficients, uint16 [n][3][64] tables, info) - no GPU involved."""
    info = probe(files[0])
    n = len(files)
    coef = torch.empty((n, info["coef_count"]), dtype=torch.int16)
    qt = torch.zeros((n, 3, 64), dtype=torch.uint16)
    _huffman(files, info, coef, qt, threads or min(n, os.cpu_count() or 1))
    return coef, qt, info


def decode_batch(files, device, threads=None):
    """n JPEG files (bytes) of one geometry -> uint8 (n, H, W, 3) RGB frames on ``device``."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.DualVarNativeError("jpeg.decode_batch decodes on a B200 (no CPU fallback)")
    info = probe(files[0])
    n = len(files)
    key = (n, info["coef_count"])
    hit = _pinned.get(key)
    if hit is None:
        _pinned.clear()          # one geometry at a time: frame batches of a run share it
        hit = _pinned[key] = (torch.empty((n, info["coef_count"]), dtype=torch.int16).pin_memory(),
                              torch.zeros((n, 3, 64), dtype=torch.uint16).pin_memory())
    coef_h, qt_h = hit
    torch.cuda.current_stream(device).synchronize()     # the previous batch's upload has left the pinned buffers
    _huffman(files, info, coef_h, qt_h, threads or min(n, os.cpu_count() or 1))
    coef = coef_h.to(device, non_blocking=True)
    qt = qt_h.to(device, non_blocking=True)
    planes = torch.empty(n * int(_lib.load().dv_jpeg_plane_bytes(info["_info"])), dtype=torch.uint8, device=device)
    out = torch.empty((n, info["height"], info["width"], 3), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.call("dv_jpeg_idct_rgb_u8", ptr(coef), ptr(qt), ptr(planes), ptr(out), n, info["_info"], info["coef_count"],
                  stream_ptr())
    return out
