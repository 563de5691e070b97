"""TEST HELPER: the oracle with the product's bf16 rounding points.

The product keeps conv outputs ``y``, activations ``z`` and their gradients as bf16 tensors and feeds the tensor cores
bf16 weights; everything between two stored tensors (accumulation, BatchNorm statistics and affine, the loss heads) is
fp32. ``emulate_bf16(model)`` returns a deep copy of an oracle model that rounds at the same places and nowhere else:

* backbone conv weights are rounded to bf16 once (gradients are then those w.r.t. the rounded weights, which is what a
  straight-through product computes),
* every backbone ``nn.Conv3d`` output and every backbone ``nn.ReLU`` output goes through ``RoundSTE``: value rounded to
  bf16 in forward, incoming gradient rounded to bf16 in backward (dy of the conv / dz of the activation),
* S3D-G's SelfGating output (scaled in place in the stored bf16 tensor) and a residual block that ends without a
  ReLU are rounded the same way; a MaxPool output needs no rounding (it selects stored values).

Against THIS yardstick the product differs only by summation order (and the ReLU-mask flips a last-bit difference in
``y`` can cause), so parameter gradients can be compared tensor by tensor with a tight bound - torch's autocast, the
previous yardstick, also rounds BatchNorm / pooling internals differently and only supported a median comparison.
"""
import copy

import torch
import torch.nn as nn


class RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def _round_hook(_mod, _inp, out):
    return RoundSTE.apply(out)


def backbones_of(model):
    """The clip encoders inside an oracle wrapper (encoder_q[0] / encoder_k[0] / .backbone) or the model itself."""
    found = []
    for name in ("encoder_q", "encoder_k"):
        enc = getattr(model, name, None)
        if enc is not None:
            found.append(enc[0])
    if getattr(model, "backbone", None) is not None:
        found.append(model.backbone)
    return found or [model]


def emulate_bf16(model):
    m = copy.deepcopy(model)
    for bb in backbones_of(m):
        for mod in bb.modules():
            if isinstance(mod, nn.Conv3d):
                with torch.no_grad():
                    mod.weight.copy_(mod.weight.to(torch.bfloat16).to(torch.float32))
                mod.register_forward_hook(_round_hook)
            elif isinstance(mod, nn.ReLU):
                mod.register_forward_hook(_round_hook)
            elif type(mod).__name__ == "SelfGating":          # the gate scales the stored bf16 slice in place
                mod.register_forward_hook(_round_hook)
            elif type(mod).__name__ == "BasicBlock2d" and not mod.use_final_relu:   # stored without a ReLU in front
                mod.register_forward_hook(_round_hook)
    return m


def round_input(x):
    """The ingest kernel stores the clips as bf16: feed both sides the representable values."""
    return x.to(torch.bfloat16).to(torch.float32)


def per_tensor_errors(ref_model, prod_model, skip_zero=1e-12):
    """[(name, ||g_prod - g_ref|| / ||g_ref||)] over parameters that received a non-zero reference gradient."""
    out = []
    for (n, pr), (_, pp) in zip(ref_model.named_parameters(), prod_model.named_parameters()):
        if pr.grad is None:
            continue
        assert pp.grad is not None and torch.isfinite(pp.grad).all(), n
        nr = pr.grad.float().norm().item()
        if nr <= skip_zero:
            continue
        out.append((n, (pp.grad.float() - pr.grad.float()).norm().item() / nr))
    return out
