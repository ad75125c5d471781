"""The timed unit: one iteration of the reference's ``train()`` body
(benchmarks/train_hiv.py:170-202 and the zinc / pcba / ppa / mnist variants):

    zero_grad -> model(data) -> task loss + lr*wd*phm_weight_regularization -> backward
    -> (data-parallel: all-reduce gradients) -> clip_grad_norm_(2.0) -> Adam step
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn.functional as F

from .nn import get_model_blocks, phm_weight_regularization
from .synthetic import Workload

BLOCKS = ("convs", "pooling", "downstream", "norms", "atomencoder", "bondencoders")


FUSED_LOSS = os.environ.get("PHC_NO_FUSED_LOSS", "") in ("", "0")      # A/B switch


def task_loss(logits: torch.Tensor, y: torch.Tensor, kind: str) -> torch.Tensor:
    if FUSED_LOSS and logits.is_cuda and kind in ("ce", "bce", "bce_masked", "l1"):
        from . import ops
        return ops.task_loss(logits, y, kind)               # loss + gradient in one launch (csrc/loss.cu)
    if kind in ("bce", "bce_masked"):
        # mean BCE over the labelled entries (train_hiv.py:174,178); written without boolean-mask
        # indexing so that no host sync is needed — same value as logits[mask] / y[mask]
        mask = ~torch.isnan(y)
        per = F.binary_cross_entropy_with_logits(logits, torch.where(mask, y, torch.zeros_like(y)).to(torch.float), reduction="none")
        return (per * mask).sum() / mask.sum()
    if kind == "l1":
        return (logits.squeeze() - y).abs().mean()              # train_zinc.py:192
    if kind == "ce":
        return F.cross_entropy(logits, y.view(-1))              # train_ppa.py / train_mnist.py:200
    raise ValueError(kind)


def make_optimizer(model, lr: float):
    """Adam over the six parameter blocks, weight_decay=0, as do_run() builds it (train_hiv.py:266-285)."""
    groups = []
    for b in BLOCKS:
        groups += get_model_blocks(model, b, lr=lr, weight_decay=0.0)
    n_groups = sum(p.numel() for g in groups for p in g["params"])
    n_model = sum(p.numel() for p in model.parameters())
    assert n_groups == n_model, "parameter blocks do not cover the model"        # train_hiv.py:279-282
    # same update rule as the scripts' torch.optim.Adam(params); on CUDA use the single-kernel-per-group
    # implementation so the step is not dominated by optimizer launches
    fused = all(p.is_cuda for g in groups for p in g["params"])
    if fused:
        # all six blocks share lr / weight_decay: one group is the same update with one fused kernel per step
        return torch.optim.Adam([p for g in groups for p in g["params"]], lr=lr, weight_decay=0.0, fused=True)
    return torch.optim.Adam(groups)


class TrainStep(object):
    """optimizer=None -> the flat two-kernel clip+Adam (optim.FlatClipAdam, same update rule); pass a torch optimizer
    (e.g. make_optimizer(model, lr)) to run the reference's torch.optim.Adam + clip_grad_norm_ instead."""

    def __init__(self, model, workload: Workload, optimizer=None, dp=None):
        from .optim import FlatClipAdam
        self.model = model
        self.wl = workload
        self.flat_opt = optimizer is None and next(model.parameters()).is_cuda
        if self.flat_opt:
            make_optimizer(model, workload.lr)     # keeps the scripts' "blocks cover all parameters" assertion
            self.opt = FlatClipAdam(model, lr=workload.lr, max_norm=workload.grad_clip)
        else:
            self.opt = optimizer if optimizer is not None else make_optimizer(model, workload.lr)
        self.dp = dp                      # DataParallelPHC wrapper or None
        if dp is not None and self.flat_opt:
            # Default: ONE all-reduce of the flat gradient buffer after backward (ncclAvg, no separate scaling pass).
            # PHC_OVERLAP_ALLREDUCE=1: slices are all-reduced as backward completes them (parallel.GradientBucket.enable_overlap).
            # Measured at 8 GPUs on the ppa workload the overlapped mode is SLOWER and noisy (84.9k - 94.5k graphs/s against
            # 97.0k): the step's persistent tcgen05 kernels need all 148 SMs, so an NCCL kernel running beside them delays
            # whichever came second on every rank, and the ring stalls on the slowest one; the buffer is only 4.7 MB.
            self.opt.bucket.group = dp.group
            if os.environ.get("PHC_OVERLAP_ALLREDUCE", "") not in ("", "0"):
                self.opt.bucket.enable_overlap(model, dp.group)
        # the scripts pick the regulariser by family (train_hiv.py:182: quaternion models have no ``phm_dim``)
        if hasattr(getattr(model, "module", model), "phm_dim"):
            self.regulariser = phm_weight_regularization
        else:
            from .quaternion import quaternion_weight_regularization
            self.regulariser = quaternion_weight_regularization
        self.params = [p for p in model.parameters()]

    def forward_backward(self, data) -> torch.Tensor:
        """zero_grad -> forward -> loss + regulariser -> backward (no collective, no optimizer): the part graphed.GraphedTrainStep
        records when the job is data parallel (the all-reduce stays an eager NCCL call between two graph launches)."""
        wl = self.wl
        self.opt.zero_grad()
        logits = self.model(data)
        loss = task_loss(logits, data.y, wl.loss)
        if wl.weight_decay > 0.0:
            loss = loss + wl.lr * wl.weight_decay * self.regulariser(self.model, p=2)
        loss.backward()
        return loss.detach()

    def optimizer_step(self) -> None:
        """[all-reduce] -> clip -> Adam on the flat buffers (flat optimizer only)."""
        self.opt.step(reduce=self.dp is not None, reduce_group=self.dp.group if self.dp is not None else None)

    def __call__(self, data) -> torch.Tensor:
        wl = self.wl
        if self.flat_opt:
            loss = self.forward_backward(data)
            self.optimizer_step()
            return loss
        self.opt.zero_grad()
        logits = self.model(data)
        loss = task_loss(logits, data.y, wl.loss)
        if wl.weight_decay > 0.0:
            loss = loss + wl.lr * wl.weight_decay * self.regulariser(self.model, p=2)
        loss.backward()
        if self.dp is not None:
            self.dp.reduce_gradients()
        if wl.grad_clip > 0.0:
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=wl.grad_clip, norm_type=2)
        self.opt.step()
        return loss.detach()
