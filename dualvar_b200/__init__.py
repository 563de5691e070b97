"""dualvar_b200 — B200-native (sm_100a) implementation of DualVar's pretraining hot path.

Host side mirrors the reference's module API (select_backbone / SimCLR_* / MoCo_* /
LinearClassifier); all arithmetic runs in hand-written CUDA kernels behind the C ABI declared in
include/dualvar_b200.h.
"""
__version__ = "0.1.0"
