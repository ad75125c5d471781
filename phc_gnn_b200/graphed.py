"""The whole training step as ONE CUDA graph, for the launch-bound workloads.

A molecule-sized step (hiv / zinc / mnist / cifar shapes: ~130 kernels, ~1.2 ms of GPU time) costs ~2.8 ms of host time in
Python, ctypes and the autograd engine — the GPU idles more than half of the step.  ``GraphedTrainStep`` records one iteration of
the reference's ``train()`` body (benchmarks/train_hiv.py:170-202: zero_grad -> forward -> loss + regulariser -> backward ->
[gradient all-reduce] -> clip -> Adam) into a ``torch.cuda.CUDAGraph`` per batch SHAPE and replays it: the host then issues a few
input copies and one graph launch per step.

What makes the step replayable (nothing the host passes by value may change between steps):
  * CSR / segment construction, all layers, loss, backward and the optimizer are kernels on the capture stream; the C ABI never
    allocates or synchronises (include/phc_b200.h);
  * Adam's step counter and learning rate live in device memory (``phc_adam_clip_step_dev``);
  * dropout masks are keyed by seed + a device-resident epoch word that the graph advances at its top
    (``phc_dropout_epoch_register`` / ``_advance``), so every replay draws fresh masks;
  * gradients are written in place into the flat gradient buffer (layer._ConvLayerDirect).
Data parallel (world size > 1): the graph holds zero_grad .. backward + the gradient pack; the all-reduce of the flat buffer stays ONE
eager NCCL call between the graph launch and the two optimizer kernels (recording NCCL collectives issued from autograd-hook threads
into a capture hung in testing, and the launch-bound workloads this path exists for have < 2 MB of gradients — nothing to overlap).
(The optional overlapped, sliced all-reduce of parallel.GradientBucket belongs to the eager step.)

A graph is specific to (N, E, B) and the feature widths.  A shape is captured when it is seen for the ``capture_after``-th time
(default: the second), at most ``max_graphs`` shapes are kept; every other call runs the eager step.  Training over a fixed set
of pre-collated batches (the benchmark, or an epoch cache) replays always; freshly shuffled batches of new shapes stay eager.
"""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import Optional

import torch

from . import _lib, graph
from .graph import _stream
from .ops import PROFILE, run

_FIELDS = ("x", "edge_index", "edge_attr", "batch", "y")
_EPOCH = {}          # device index -> the registered dropout epoch word


def dropout_epoch(device: torch.device) -> torch.Tensor:
    """The process-wide dropout epoch word of this device (registered with the library on first use)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    t = _EPOCH.get(idx)
    if t is None:
        t = torch.zeros(1, dtype=torch.int64, device=device)
        _lib.check(_lib.load().phc_dropout_epoch_register(t.data_ptr()), "phc_dropout_epoch_register")
        _EPOCH[idx] = t
    return t


class _Entry(object):
    __slots__ = ("graph", "static", "loss", "launches")


class GraphedTrainStep(object):
    def __init__(self, step, max_graphs: int = 32, capture_after: int = 2):
        """``step``: a train.TrainStep built with the flat optimizer (optimizer=None)."""
        assert step.flat_opt, "GraphedTrainStep needs the flat clip+Adam optimizer (its step counter lives on the device)"
        self.step = step
        self.split = step.dp is not None                    # data parallel: collective + optimizer stay outside the graph
        if self.split:
            step.opt.bucket.overlap = False                 # one all-reduce after the graph; no NCCL calls from backward hooks
        self.max_graphs, self.capture_after = int(max_graphs), int(capture_after)
        self.entries: "OrderedDict[tuple, _Entry]" = OrderedDict()
        self.seen = {}
        self.pool = None
        self.replays = self.eager_steps = self.captures = 0

    @staticmethod
    def key(data) -> tuple:
        return tuple((tuple(getattr(data, f).shape), getattr(data, f).dtype) for f in _FIELDS) + (int(data.num_graphs),)

    def __call__(self, data) -> torch.Tensor:
        if PROFILE.timing or not data.x.is_cuda:            # per-operator CUDA-event timing needs the individual calls
            self.eager_steps += 1
            return self.step(data)
        k = self.key(data)
        ent = self.entries.get(k)
        if ent is None:
            n = self.seen.get(k, 0) + 1
            self.seen[k] = n
            if n < self.capture_after or len(self.entries) >= self.max_graphs:
                self.eager_steps += 1
                return self.step(data)
            return self._capture(k, data)
        self.entries.move_to_end(k)
        for f in _FIELDS:
            getattr(ent.static, f).copy_(getattr(data, f), non_blocking=True)
        self.step.opt.sync_lr()
        ent.graph.replay()
        PROFILE.launches += ent.launches
        if self.split:
            self.step.optimizer_step()
        self.replays += 1
        return ent.loss

    def _capture(self, k, data) -> torch.Tensor:
        """Runs this call's step eagerly on the static input buffers (it is a real training step and the warm-up of every lazily
        created buffer), then records the same step into a graph WITHOUT executing it."""
        dev = data.x.device
        dropout_epoch(dev)                                  # registered before the first captured dropout kernel is parameterised
        static = copy.copy(data)
        for f in _FIELDS:
            setattr(static, f, getattr(data, f).clone())
        loss_now = self.step(static)                        # the step this call owes (eager)
        self.step.opt.sync_lr()
        ent = _Entry()
        ent.static = static
        ent.graph = torch.cuda.CUDAGraph()
        graph.clear_cache()
        before = PROFILE.launches
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        with torch.cuda.graph(ent.graph, pool=self.pool):
            ep = dropout_epoch(dev)
            run("phc_dropout_epoch_advance", None, ep.data_ptr(), _stream(dev))
            graph.clear_cache()                             # the CSR / segment build is part of every step
            if self.split:
                ent.loss = self.step.forward_backward(static)
                self.step.opt.bucket.pack()                 # every gradient in its slice of the flat buffer, inside the graph
            else:
                ent.loss = self.step(static)
        ent.launches = PROFILE.launches - before
        PROFILE.launches = before
        graph.clear_cache()                                 # structures built during capture live in the graph's pool
        self.entries[k] = ent
        self.captures += 1
        return loss_now

    def stats(self) -> dict:
        return dict(graphs=len(self.entries), captures=self.captures, replays=self.replays, eager_steps=self.eager_steps)
