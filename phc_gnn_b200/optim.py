"""Flat-buffer optimizer: gradient clipping by global norm + Adam in two kernel launches per step.

Same update as the reference's ``clip_grad_norm_(model.parameters(), 2.0)`` followed by
``torch.optim.Adam(..., weight_decay=0).step()`` (benchmarks/train_hiv.py:199-201, :266-285), but all
parameters live in one flat fp32 buffer (each ``nn.Parameter`` is re-pointed at its slice, so modules,
state-dicts and the reference's ``get_model_blocks`` still see ordinary parameters), gradients are packed
into one flat buffer (the same buffer the data-parallel all-reduce uses), and csrc/optimizer.cu does the
norm, the clipping and the Adam update.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.nn as nn

from . import _lib
from .flat import alias_flat
from .graph import _stream
from .ops import _ws, run
from .parallel import GradientBucket


def ordered_parameters(model: nn.Module) -> List[nn.Parameter]:
    """All trainable parameters; the per-component BatchNorm weights (then biases) of every NaivePHMNorm are kept
    adjacent so that the norm kernels' flat gamma / beta vectors are slices of the optimizer's flat buffer."""
    from .nn import NaivePHMNorm
    seen, out = set(), []
    for m in model.modules():
        if isinstance(m, NaivePHMNorm) and m.affine:
            for p in [b.weight for b in m.bn] + [b.bias for b in m.bn]:
                if id(p) not in seen and p.requires_grad:
                    seen.add(id(p))
                    out.append(p)
    rest = [p for p in model.parameters() if id(p) not in seen and p.requires_grad]
    return rest + out


class FlatClipAdam(object):
    def __init__(self, model_or_params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 2.0):
        if isinstance(model_or_params, nn.Module):
            self.params = ordered_parameters(model_or_params)
        else:
            self.params = [p for p in model_or_params if p.requires_grad]
        self.lr, self.betas, self.eps, self.max_norm = float(lr), betas, float(eps), float(max_norm)
        self.bucket = GradientBucket(self.params)
        self.flat = None
        self.exp_avg = self.exp_avg_sq = None
        self.t = 0
        self.grad_norm = None
        self.param_groups = [dict(params=self.params, lr=self.lr)]     # so lr schedulers can drive it

    def _ensure(self):
        self.flat = alias_flat(self.flat, self.params)
        if self.exp_avg is None or self.exp_avg.device != self.flat.device:
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
            self.grad_norm = torch.zeros((), dtype=torch.float32, device=self.flat.device)
            self._ws = _ws(_lib.load().phc_adam_workspace_bytes(), self.flat.device)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    def state_dict(self):
        return dict(t=self.t, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, lr=self.param_groups[0]["lr"])

    @torch.no_grad()
    def step(self, reduce_group=None, reduce: bool = False):
        """Pack gradients (and all-reduce them when data parallel), clip by global norm, Adam update."""
        self._ensure()
        if reduce:
            self.bucket.reduce(reduce_group)
            g = self.bucket.flat
        else:
            g = self.bucket.pack()
        self.t += 1
        b1, b2 = self.betas
        lr = float(self.param_groups[0]["lr"])
        dev = self.flat.device
        run("phc_adam_clip_step", dev, self.flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.flat.numel(), lr, b1, b2, self.eps, 1.0 - b1 ** self.t, 1.0 - b2 ** self.t, self.max_norm, self.grad_norm.data_ptr(),
            self._ws.data_ptr(), self._ws.numel(), _stream(dev))
