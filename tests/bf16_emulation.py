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


def round_conv_weights(model):
    """In place: backbone conv weights to bf16-representable values (what the tensor cores are fed). The fp32 oracle, its
    emulating copy and the product then all start from identical weights."""
    for bb in backbones_of(model):
        for mod in bb.modules():
            if isinstance(mod, nn.Conv3d):
                with torch.no_grad():
                    mod.weight.copy_(mod.weight.to(torch.bfloat16).to(torch.float32))
    return model


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


def per_tensor_errors(ref_model, prod_model, skip_rel=1e-5):
    """[(name, ||g_prod - g_ref|| / ||g_ref||)] over parameters that received a reference gradient. Tensors whose
    reference gradient is numerically zero (a conv bias in front of training-mode BatchNorm: fp32 round-off only, norm
    below skip_rel of the largest gradient) are left out."""
    norms = [p.grad.float().norm().item() for p in ref_model.parameters() if p.grad is not None]
    floor = skip_rel * max(norms) if norms else 0.0
    out = []
    for (n, pr), (_, pp) in zip(ref_model.named_parameters(), prod_model.named_parameters()):
        if pr.grad is None:
            continue
        assert pp.grad is not None and torch.isfinite(pp.grad).all(), n
        nr = pr.grad.float().norm().item()
        if nr <= floor:
            continue
        out.append((n, (pp.grad.float() - pr.grad.float()).norm().item() / nr))
    return out


def compare_to_noise_floor(ref_model, emu_model, prod_model):
    """Tensor by tensor: the product's gradient error against the fp32 oracle next to the error the SAME rounding
    points cause in the oracle itself (emu vs fp32) - bf16 gradients through dozens of training-mode BatchNorm + ReLU
    layers carry 10-40 % L2 noise (ReLU-mask flips of last-bit differences), so the bound on a tensor is its own noise
    floor, not a constant. Returns [(name, err_prod, err_emu)]."""
    ep = dict(per_tensor_errors(ref_model, prod_model))
    ee = dict(per_tensor_errors(ref_model, emu_model))
    return [(n, ep[n], ee[n]) for n in ep if n in ee]
