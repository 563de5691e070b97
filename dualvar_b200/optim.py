"""Multi-tensor SGD(momentum, weight_decay) on one kernel launch (csrc/moco.cu: sgd_momentum_kernel).

Mirrors ``torch.optim.SGD(params, lr, momentum, weight_decay)`` as the reference uses it
(pretrain.py:262-272: one param group per tensor, dampening 0, no nesterov): same update rule, same
``param_groups`` interface (``lr`` can be changed per group by a scheduler such as MultiStepLR,
pretrain.py:328), ``zero_grad`` and ``state_dict``. Groups sharing (lr, momentum, weight_decay) are
batched into a single launch.
"""
import ctypes

import torch

from . import _lib
from ._lib import ptr, stream_ptr

_CHUNK = 8192


class SGD(torch.optim.Optimizer):
    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, dampening=0.0, nesterov=False, maximize=False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        # the full torch.optim.SGD group layout, so state_dict() loads into torch.optim.SGD and back
        # (checkpoint interchange with the reference, pretrain.py:262-272, :301); the kernel implements the
        # reference's setting of the three extra knobs only
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov, maximize=maximize, foreach=None, differentiable=False,
                                      fused=None))
        self._tables = {}

    @staticmethod
    def _check_group(group):
        if group.get("dampening", 0.0) != 0.0 or group.get("nesterov", False) or group.get("maximize", False):
            raise _lib.DualVarNativeError("dualvar_b200.optim.SGD implements dampening=0, nesterov=False, maximize=False "
                                          "(the reference's optimizer, pretrain.py:272)")

    def _table(self, key, plist):
        sig = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["momentum_buffer"].data_ptr(), p.numel())
                    for p in plist)
        hit = self._tables.get(key)
        if hit is None or hit[0] != sig:
            rows = []
            for pp, gp, bp, n in sig:
                for off in range(0, n, _CHUNK):
                    rows.append((pp + 4 * off, gp + 4 * off, bp + 4 * off, min(_CHUNK, n - off)))
            hit = (sig, torch.tensor(rows, dtype=torch.int64).to(plist[0].device))
            self._tables[key] = hit
        return hit[1]

    def zero_grad(self, set_to_none=True):
        """As torch.optim.Optimizer.zero_grad; the in-place form zeroes all gradients with multi-tensor launches instead
        of one fill kernel per parameter (one param group per tensor, pretrain.py:262-264, defeats torch's grouping)."""
        if set_to_none:
            return super().zero_grad(set_to_none=True)
        grads = [p.grad for group in self.param_groups for p in group["params"] if p.grad is not None]
        for g in grads:
            if g.grad_fn is not None:
                g.detach_()
            else:
                g.requires_grad_(False)
        if grads:
            torch._foreach_zero_(grads)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        batches = {}
        for group in self.param_groups:
            self._check_group(group)
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise _lib.DualVarNativeError("dualvar_b200.optim.SGD needs contiguous fp32 CUDA parameters")
                st = self.state[p]
                first = "momentum_buffer" not in st
                if first:
                    st["momentum_buffer"] = torch.empty_like(p)
                key = (group["lr"], group["momentum"], group["weight_decay"], first)
                batches.setdefault(key, []).append(p)
        for (lr, mu, wd, first), plist in batches.items():
            t = self._table((len(plist), plist[0].data_ptr()), plist)      # pointers only: survives an lr change and the first step
            f = ctypes.c_float
            _lib.call("dv_sgd_momentum_step", ptr(t), t.shape[0], f(lr), f(mu), f(wd), 1 if first else 0, stream_ptr())
        from . import engine
        engine.invalidate_weights(p for ps in batches.values() for p in ps)
        return loss
