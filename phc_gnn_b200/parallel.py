"""Data-parallel training across the GPUs of one box: one process per GPU, full model replica per
rank, graph mini-batches sharded by rank, ONE all-reduce per step over a flat fp32 gradient buffer
(NCCL over NVLink 5 / NVSwitch; gloo on CPU for tests).

The reference has no distributed path at all; the only hook is
``get_model_blocks`` tolerating a ``.module`` wrapper (phc/quaternion/regularization.py:6-8), which
this wrapper satisfies.  Message passing, CSR construction and pooling never cross ranks; BatchNorm
uses per-rank batch statistics (DDP semantics).
"""
from __future__ import annotations

import weakref
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


# parameter -> its slice of a flat gradient buffer (set by GradientBucket, read by layer.py's in-place gradient path);
# a side table rather than an attribute so that it is never pickled with the parameter
_SINKS = {}     # id(param) -> (weakref to the parameter, view); keyed by id because tensors do not compare as scalars


def register_grad_sink(param: torch.nn.Parameter, view: torch.Tensor) -> None:
    key = id(param)
    _SINKS[key] = (weakref.ref(param, lambda _r, k=key: _SINKS.pop(k, None)), view)


def grad_sink(param) -> Optional[torch.Tensor]:
    e = _SINKS.get(id(param))
    return e[1] if (e is not None and e[0]() is param) else None


class GradientBucket(object):
    """Flat fp32 buffer that all gradients are packed into after backward; after ``reduce()`` every
    ``param.grad`` is a view of the (averaged) flat buffer, so clipping and the optimizer step read
    the reduced values without a copy back."""

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None
        self.views: List[torch.Tensor] = []

    def _ensure(self, device):
        if self.flat is None or self.flat.device != device:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
            for p, v in zip(self.params, self.views):
                register_grad_sink(p, v)   # layer.py writes this parameter's gradient straight into the flat buffer

    def pack(self):
        dev = self.params[0].device
        self._ensure(dev)
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is v:
                continue                   # written in place by the layer's backward
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                src.append(g)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            p.grad = v
        return self.flat

    def reduce(self, group=None, async_op: bool = False):
        flat = self.pack()
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return None
        if dist.get_backend(group) == "nccl":
            # ncclAvg: the 1/world factor is applied inside the collective (no separate pass over the buffer)
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        flat.div_(world)        # gloo (CPU tests) has no AVG
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class DataParallelPHC(nn.Module):
    """Thin replica wrapper: broadcasts rank 0's parameters/buffers at construction and exposes
    ``reduce_gradients()`` to be called between ``backward()`` and ``clip_grad_norm_`` / ``step()``
    (the order of benchmarks/train_hiv.py:198-201 on the reduced gradients)."""

    def __init__(self, module: nn.Module, group=None):
        super().__init__()
        self.module = module
        self.group = group
        self.bucket = GradientBucket(list(module.parameters()))
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=0, group=group)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def reduce_gradients(self, async_op: bool = False):
        return self.bucket.reduce(self.group, async_op)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)
