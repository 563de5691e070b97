"""TEST HELPERS (not product code): Python-side handles on single C-ABI kernels for the unit tests and tests/diag -
allocate outputs with torch, pass raw pointers. The product path calls the C ABI from dualvar_b200/engine.py,
objectives.py and frames.py directly; nothing under dualvar_b200/ imports this module.

torch is used for device memory and streams only; every arithmetic op below is a kernel from
dualvar_b200/csrc launched through the C ABI (include/dualvar_b200.h).
"""
import ctypes

import torch

from dualvar_b200 import _lib
from dualvar_b200._lib import ConvGeom, make_geom, pad8, ptr, stream_ptr  # noqa: F401


def _require_cuda(t, name):
    if not t.is_cuda:
        raise _lib.DualVarNativeError(
            f"{name}: tensor is on {t.device}; dualvar_b200 kernels run on a B200 only (no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.DualVarNativeError(f"{name}: tensor must be contiguous")


# ----------------------------------------------------------------------------- layout
def to_ndhwc(x):
    """fp32 NCDHW -> bf16 NDHWC with channels padded to a multiple of 8."""
    _require_cuda(x, "to_ndhwc")
    N, C, T, H, W = x.shape
    Cp = pad8(C)
    y = torch.empty((N, T, H, W, Cp), dtype=torch.bfloat16, device=x.device)
    _lib.call("dv_ncdhw_to_ndhwc_bf16", ptr(x), ptr(y), N, C, Cp, T * H * W, stream_ptr())
    return y


def from_ndhwc(y, C):
    """bf16 NDHWC (padded) -> fp32 NCDHW with C logical channels."""
    _require_cuda(y, "from_ndhwc")
    N, T, H, W, Cp = y.shape
    x = torch.empty((N, C, T, H, W), dtype=torch.float32, device=y.device)
    _lib.call("dv_ndhwc_bf16_to_ncdhw", ptr(y), ptr(x), N, C, Cp, T * H * W, stream_ptr())
    return x


def pack_conv_weight(w, g, want_fprop=True, want_dgrad=True):
    """nn.Conv3d weight (fp32, Cout x Cin x kt x kh x kw) -> packed bf16 operands."""
    _require_cuda(w, "pack_conv_weight")
    wf = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.bfloat16, device=w.device) if want_fprop else None
    wt = torch.empty((g.Cin_p, g.taps, g.Cout_p), dtype=torch.bfloat16, device=w.device) if want_dgrad else None
    _lib.call("dv_pack_conv_weight", ptr(w), ptr(wf), ptr(wt), ctypes.byref(g), stream_ptr())
    return wf, wt


# ----------------------------------------------------------------------------- convolution
def conv3d_fprop(x, wf, g, bn_stats=None, bias=None):
    _require_cuda(x, "conv3d_fprop")
    y = torch.empty((g.N, g.To, g.Ho, g.Wo, g.Cout_p), dtype=torch.bfloat16, device=x.device)
    _lib.call("dv_conv3d_fprop_bf16", ptr(x), ptr(wf), ptr(y), ptr(bn_stats), ptr(bias),
              ctypes.byref(g), stream_ptr())
    return y


def conv3d_dgrad(dy, wt, g):
    _require_cuda(dy, "conv3d_dgrad")
    dx = torch.empty((g.N, g.T, g.H, g.W, g.Cin_p), dtype=torch.bfloat16, device=dy.device)
    _lib.call("dv_conv3d_dgrad_bf16", ptr(dy), ptr(wt), ptr(dx), ctypes.byref(g), stream_ptr())
    return dx


def conv3d_wgrad_packed(x, dy, g):
    _require_cuda(x, "conv3d_wgrad")
    dwp = torch.empty((g.Cout_p, g.taps, g.Cin_p), dtype=torch.float32, device=x.device)
    _lib.call("dv_conv3d_wgrad_bf16", ptr(x), ptr(dy), ptr(dwp), ctypes.byref(g), stream_ptr())
    return dwp


def unpack_conv_wgrad(dwp, g, grad=None, beta=0.0):
    if grad is None:
        grad = torch.empty((g.Cout, g.Cin, g.kt, g.kh, g.kw), dtype=torch.float32, device=dwp.device)
        beta = 0.0
    _lib.call("dv_unpack_conv_wgrad", ptr(dwp), ptr(grad), ctypes.byref(g), ctypes.c_float(beta), stream_ptr())
    return grad
