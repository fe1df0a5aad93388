"""Optimisers for the flat parameter buffer (reference: utils/util_methods.py:14-19).

`get_optimizer(Model, opt)` keeps the reference's signature and hyper-parameters
(SGD: momentum 0.9, weight_decay 1e-6; Adam: betas (0.9, 0.98); lr = opt.lr) and returns a
`torch.optim.Optimizer` subclass whose `step()` is one CUDA kernel over the model's flat
parameter / gradient buffers, so `StepLR` (main.py:86) and `zero_grad()` keep working.
A stock `torch.optim.SGD` / `Adam` on the same parameters works too.
"""
from __future__ import annotations

import torch

from . import ops
from .engine import flat_params


class _FlatOptimizer(torch.optim.Optimizer):
    def __init__(self, model, defaults):
        self.model = model
        super().__init__([p for p in model.parameters()], defaults)
        self._flat_buffers = {}

    def _buffer(self, name, like):
        b = self._flat_buffers.get(name)
        if b is None or b.shape != like.shape or b.device != like.device:
            b = torch.zeros_like(like)
            self._flat_buffers[name] = b
        return b

    def zero_grad(self, set_to_none: bool = True):
        """Always clears the flat gradient buffer (one ~190 KB memset).  Every `p.grad` is a view into it, so
        `set_to_none` cannot detach them without breaking the one-kernel step; zeroing is what makes the
        reference's own loop (`zero_grad(); loss.backward(); step()`, finetune.py:39-49) correct when the
        gradients come from autograd (`_ChromeGCNFn`), whose AccumulateGrad nodes add into `p.grad` in place.
        The fused `cgcn_train_step` overwrites the buffer anyway."""
        fp = flat_params(self.model, full=False)
        fp.flat_grad.zero_()
        fp.attach_grads()

    def state_dict(self):
        sd = super().state_dict()
        sd["flat_buffers"] = {k: v.clone() for k, v in self._flat_buffers.items()}
        sd["flat_steps"] = getattr(self, "_steps", 0)
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        self._flat_buffers = {k: v.clone() for k, v in sd.pop("flat_buffers", {}).items()}
        self._steps = sd.pop("flat_steps", 0)
        super().load_state_dict(sd)


class FlatSGD(_FlatOptimizer):
    def __init__(self, model, lr, momentum=0.9, weight_decay=1e-6, grad_scale=1.0):
        super().__init__(model, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.grad_scale = grad_scale

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        fp = flat_params(self.model, full=False)
        buf = self._buffer("momentum", fp.flat)
        ops.sgd_step(fp.flat, fp.flat_grad, buf, g["lr"], g["momentum"], g["weight_decay"], self.grad_scale)


class FlatAdam(_FlatOptimizer):
    def __init__(self, model, lr, betas=(0.9, 0.98), eps=1e-8, grad_scale=1.0):
        super().__init__(model, dict(lr=lr, betas=betas, eps=eps))
        self.grad_scale = grad_scale
        self._steps = 0

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        fp = flat_params(self.model, full=False)
        m = self._buffer("exp_avg", fp.flat)
        v = self._buffer("exp_avg_sq", fp.flat)
        self._steps += 1
        ops.adam_step(fp.flat, fp.flat_grad, m, v, g["lr"], self._steps, g["betas"][0], g["betas"][1], g["eps"],
                      self.grad_scale)


def get_optimizer(Model, opt):
    """utils/util_methods.py:14-19: built from `opt.optim` / `opt.lr` (not -optim2 / -lr2)."""
    if opt.optim == "adam":
        return FlatAdam(Model, lr=opt.lr, betas=(0.9, 0.98))
    if opt.optim == "sgd":
        return FlatSGD(Model, lr=opt.lr, weight_decay=1e-6, momentum=0.9)
    raise ValueError("opt.optim must be 'adam' or 'sgd'")
