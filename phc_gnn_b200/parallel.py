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
_OVERLAP_IN_USE = False


def overlap_in_use() -> bool:
    """True once a bucket all-reduces gradient slices while backward still runs (ops._WeightReg then accumulates through autograd)."""
    return _OVERLAP_IN_USE



def register_grad_sink(param: torch.nn.Parameter, view: torch.Tensor) -> None:
    key = id(param)
    _SINKS[key] = (weakref.ref(param, lambda _r, k=key: _SINKS.pop(k, None)), view)


def grad_sink(param) -> Optional[torch.Tensor]:
    e = _SINKS.get(id(param))
    return e[1] if (e is not None and e[0]() is param) else None


class GradientBucket(object):
    """Flat fp32 buffer that all gradients are packed into; after ``reduce()`` every ``param.grad`` is a view of the (averaged)
    flat buffer, so clipping and the optimizer step read the reduced values without a copy back.

    Overlap with backward: ``params`` is ordered by gradient completion and ``stage_ends`` marks the stages
    (optim.staged_parameters).  ``enable_overlap(model, group)`` installs the model's ``_stage_hook``; when backward reports stages
    <= k final, their slice of the buffer is packed and all-reduced asynchronously (NCCL's own stream, ordered after the kernels
    issued so far) while the layers below keep running; slices smaller than ``min_chunk_bytes`` are merged with the next stage.
    ``reduce()`` after backward sends the remainder and makes the compute stream wait for everything.  The per-rank values and
    the reduction order inside each slice are the same as for the single all-reduce, so replicas stay bit-identical."""

    def __init__(self, params: List[torch.nn.Parameter], stage_ends: Optional[List[int]] = None, min_chunk_bytes: int = 1 << 20):
        keep = [p.requires_grad for p in params]
        self.params = [p for p, k in zip(params, keep) if k]
        assert all(keep) or stage_ends is None, "stage boundaries refer to the trainable parameters"
        self.stage_ends = list(stage_ends) if stage_ends else [len(self.params)]
        self.min_chunk_bytes = int(min_chunk_bytes)
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None
        self.views: List[torch.Tensor] = []
        self.offsets: List[int] = []
        self.group = None
        self.overlap = False
        self._done = 0              # parameters [0, _done) are packed and their all-reduce is in flight
        self._works = []
        self.launched = 0           # all-reduce calls of the current step (diagnostics)
        self.local_weight = 1.0     # this rank's share of the global batch relative to an even split (see set_batch_share)

    def _ensure(self, device):
        if self.flat is None or self.flat.device != device:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
            self.views, self.offsets, off = [], [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view(p.shape))
                self.offsets.append(off)
                off += p.numel()
            self.offsets.append(off)
            for p, v in zip(self.params, self.views):
                register_grad_sink(p, v)   # layer.py writes this parameter's gradient straight into the flat buffer

    def _pack_range(self, i0: int, i1: int):
        src, dst = [], []
        for i in range(i0, i1):
            p, v = self.params[i], self.views[i]
            g = p.grad
            if g is v:
                continue                   # written in place by the layer's backward
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                src.append(g)
                dst.append(v)
            p.grad = v
        if src:
            torch._foreach_copy_(dst, src)

    def pack(self):
        dev = self.params[0].device
        self._ensure(dev)
        self._pack_range(self._done, len(self.params))
        return self.flat

    # ---- overlapped reduction ------------------------------------------------------------------------
    def enable_overlap(self, model: nn.Module, group=None) -> bool:
        """Ask ``model`` (the PHC skip-connect models call ``_stage_hook`` from tensor hooks on the layer outputs) to report
        completed stages.  Returns False — and stays with the single all-reduce — for a model without stages."""
        root = getattr(model, "module", model)
        self.group = group
        if len(self.stage_ends) < 2 or not hasattr(root, "convs"):
            return False
        object.__setattr__(root, "_stage_hook", self.stage_ready)
        self.overlap = True
        global _OVERLAP_IN_USE
        _OVERLAP_IN_USE = True            # slices leave for the all-reduce DURING backward: nothing may be added to them at its end
        return True

    def set_batch_share(self, local_graphs: int, global_graphs: int, world: int) -> None:
        """Uneven shards (a short last global batch gives some ranks one graph more): every rank's mean-loss gradient is
        weighted by local_graphs * world / global_graphs before the average, so that the reduced gradient is the gradient of
        the mean over the GLOBAL batch, not a mean of per-rank means.  An even split has weight 1 (no extra pass)."""
        self.local_weight = float(local_graphs) * float(world) / float(global_graphs)

    def _all_reduce(self, i0: int, i1: int, async_op: bool):
        a, b = self.offsets[i0], self.offsets[i1]
        if b == a:
            return None
        view = self.flat[a:b]
        if self.local_weight != 1.0:
            view.mul_(self.local_weight)
        self.launched += 1
        if dist.get_backend(self.group) == "nccl":
            # ncclAvg: the 1/world factor is applied inside the collective (no separate pass over the buffer)
            return dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        view.div_(dist.get_world_size(self.group))        # gloo (CPU tests) has no AVG
        return dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def stage_ready(self, k: int) -> None:
        """Backward finished every gradient of stages <= k (called from a tensor hook, on the backward stream)."""
        if not self.overlap or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        end = self.stage_ends[min(k, len(self.stage_ends) - 1)]
        if end <= self._done or end >= len(self.params):
            return                                          # the last stage goes out with reduce()
        dev = self.params[0].device
        self._ensure(dev)
        if (self.offsets[end] - self.offsets[self._done]) * 4 < self.min_chunk_bytes:
            return                                          # too small for a launch of its own: rides with the next stage
        self._pack_range(self._done, end)
        w = self._all_reduce(self._done, end, async_op=True)
        if w is not None:
            self._works.append(w)
        self._done = end

    def reduce(self, group=None, async_op: bool = False):
        """All-reduce (average) whatever has not been sent yet and wait for every reduction of this step."""
        if group is not None:
            self.group = group
        flat = self.pack()
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1:
            w = self._all_reduce(self._done, len(self.params), async_op=bool(self._works) or async_op)
            if w is not None:
                self._works.append(w)
            if not async_op:
                for w in self._works:
                    if w is not None:
                        w.wait()                            # the compute stream waits for NCCL's stream
                self._works = []
        self._done = 0
        self.launched_last, self.launched = self.launched, 0
        return flat


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
        self.bucket.reduce(self.group, async_op)
        return None

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)
