"""dualvar_b200 — B200-native (sm_100a) implementation of DualVar's pretraining hot path.

Host side mirrors the reference's module API (select_backbone / SimCLR_* / MoCo_* /
LinearClassifier); all arithmetic runs in hand-written CUDA kernels behind the C ABI declared in
include/dualvar_b200.h.
"""
__version__ = "0.1.0"


def set_precision(mode, planes=None):
    """"bf16" (default, throughput mode) or "fp32" (the 1e-4 parity mode: fp32 activations, convolutions as sums of
    bf16 split-plane products on the tcgen05 kernels; planes = 3 (24 mantissa bits, default) or 2). See engine.py."""
    from . import engine
    engine.set_precision(mode, planes)
