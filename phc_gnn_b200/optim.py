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


def _norm_first(module: nn.Module, seen: set) -> List[nn.Parameter]:
    """Trainable parameters of ``module``: first the per-component BatchNorm weights (then biases) of every NaivePHMNorm inside it,
    kept adjacent so that the norm kernels' flat gamma / beta vectors are slices of the optimizer's flat buffer, then the rest."""
    from .nn import NaivePHMNorm
    out = []
    for m in module.modules():
        if isinstance(m, NaivePHMNorm) and m.affine:
            for p in [b.weight for b in m.bn] + [b.bias for b in m.bn]:
                if id(p) not in seen and p.requires_grad:
                    seen.add(id(p))
                    out.append(p)
    for p in module.parameters():
        if id(p) not in seen and p.requires_grad:
            seen.add(id(p))
            out.append(p)
    return out


def staged_parameters(model: nn.Module):
    """(params, ends): all trainable parameters in the order in which backward COMPLETES their gradients, and the running
    count after each stage.  For the PHC message-passing models (nn._PHMSkipConnectBase): stage 0 = pooling + downstream head,
    stage k = message-passing layer L-k (its conv, norm and bond encoder), last stage = layer 0 + the atom encoder.  The model
    fires ``_stage_hook(k)`` when the gradient of layer L-1-k's output is complete, i.e. when stages <= k are final — which lets
    parallel.GradientBucket all-reduce them while the layers below still run.  Any other module is a single stage."""
    root = getattr(model, "module", model)
    seen: set = set()
    params: List[nn.Parameter] = []
    ends: List[int] = []
    convs, norms, bonds = getattr(root, "convs", None), getattr(root, "norms", None), getattr(root, "bondencoders", None)
    if isinstance(convs, nn.ModuleList) and hasattr(root, "pooling") and hasattr(root, "downstream") and len(convs) > 0:
        L = len(convs)
        params += _norm_first(root.pooling, seen) + _norm_first(root.downstream, seen)
        ends.append(len(params))
        for i in range(L - 1, -1, -1):
            for group in (convs, norms, bonds):
                if isinstance(group, nn.ModuleList) and i < len(group) and isinstance(group[i], nn.Module):
                    params += _norm_first(group[i], seen)
            if i > 0:
                ends.append(len(params))
    params += _norm_first(root, seen)        # atom encoder and anything not covered above
    ends.append(len(params))
    return params, ends


def ordered_parameters(model: nn.Module) -> List[nn.Parameter]:
    return staged_parameters(model)[0]


class FlatClipAdam(torch.optim.Optimizer):
    """A ``torch.optim.Optimizer`` (lr schedulers such as the scripts' ReduceLROnPlateau / StepLR, train_hiv.py:287, drive it through
    ``param_groups[0]["lr"]``) whose whole update is two kernel launches on flat buffers.  The learning rate and the step counter
    live in device memory (``lr_dev``, ``step_dev``), so ``step()`` passes no host scalar that changes between steps and can be
    recorded into a CUDA graph (graphed.GraphedTrainStep).  Parameters whose ``.grad`` is None contribute a zero gradient (their
    Adam moments still decay and move them; torch.optim.Adam would skip them) — every parameter of the PHC models gets a gradient
    each step, see ``strict_grads`` to make a missing one an error."""

    def __init__(self, model_or_params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 2.0,
                 strict_grads: bool = False):
        stage_ends = None
        if isinstance(model_or_params, nn.Module):
            params, stage_ends = staged_parameters(model_or_params)
        else:
            params = [p for p in model_or_params if p.requires_grad]
        super().__init__(params, dict(lr=float(lr), betas=tuple(betas), eps=float(eps), max_norm=float(max_norm)))
        self.params = params
        self.lr, self.betas, self.eps, self.max_norm = float(lr), tuple(betas), float(eps), float(max_norm)
        self.strict_grads = strict_grads
        self.bucket = GradientBucket(self.params, stage_ends)
        self.flat = None
        self.exp_avg = self.exp_avg_sq = None
        self.grad_norm = None
        self.lr_dev = self.step_dev = None
        self._lr_uploaded = None
        self._t_pending = 0             # steps restored by load_state_dict before the device buffers exist

    @property
    def t(self) -> int:
        """Number of steps taken (reads the device counter: synchronises)."""
        return int(self.step_dev.item()) if self.step_dev is not None else self._t_pending

    def _ensure(self):
        self.flat = alias_flat(self.flat, self.params)
        if self.exp_avg is None or self.exp_avg.device != self.flat.device:
            dev = self.flat.device
            old = (self.exp_avg, self.exp_avg_sq, self.t)
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
            if old[0] is not None and old[0].numel() == self.flat.numel():      # moved to another device: keep the state
                self.exp_avg.copy_(old[0])
                self.exp_avg_sq.copy_(old[1])
            self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)
            self.lr_dev = torch.zeros((), dtype=torch.float32, device=dev)
            self.step_dev = torch.full((), old[2], dtype=torch.int32, device=dev)
            self._lr_uploaded = None
            self._ws = _ws(_lib.load().phc_adam_workspace_bytes(), dev)

    def sync_lr(self):
        """Upload the learning rate when a scheduler changed ``param_groups[0]['lr']`` (a device write, no sync)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_uploaded:
            self.lr_dev.fill_(lr)
            self._lr_uploaded = lr

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    def state_dict(self):
        return dict(t=self.t, exp_avg=None if self.exp_avg is None else self.exp_avg.detach().clone(),
                    exp_avg_sq=None if self.exp_avg_sq is None else self.exp_avg_sq.detach().clone(), lr=float(self.param_groups[0]["lr"]),
                    betas=self.betas, eps=self.eps, max_norm=self.max_norm)

    def load_state_dict(self, state):
        """Resume: moments, step count and learning rate as ``state_dict()`` wrote them (the flat layout is the order of
        ``ordered_parameters`` — the same model class gives the same layout)."""
        self.param_groups[0]["lr"] = float(state["lr"])
        self.betas, self.eps, self.max_norm = tuple(state.get("betas", self.betas)), float(state.get("eps", self.eps)), float(state.get("max_norm", self.max_norm))
        if state.get("exp_avg") is None:
            self._t_pending = int(state["t"])
            return
        if next(iter(self.params)).is_cuda:
            self._ensure()
            assert state["exp_avg"].numel() == self.flat.numel(), "optimizer state belongs to a different parameter layout"
            self.exp_avg.copy_(state["exp_avg"])
            self.exp_avg_sq.copy_(state["exp_avg_sq"])
            self.step_dev.fill_(int(state["t"]))
        else:
            raise RuntimeError("FlatClipAdam.load_state_dict: move the model to its CUDA device first")

    @torch.no_grad()
    def step(self, closure=None, reduce_group=None, reduce: bool = False):
        """Pack gradients (and all-reduce them when data parallel), clip by global norm, Adam update."""
        assert closure is None, "FlatClipAdam does not re-evaluate a closure"
        self._ensure()
        if self.strict_grads:
            missing = [i for i, p in enumerate(self.params) if p.grad is None]
            if missing:
                raise RuntimeError(f"FlatClipAdam(strict_grads=True): {len(missing)} parameters have no gradient")
        if reduce:
            self.bucket.reduce(reduce_group)          # whatever the overlapped stage reductions have not covered yet, then waits for all
            g = self.bucket.flat
        else:
            g = self.bucket.pack()
        self.sync_lr()
        b1, b2 = self.betas
        dev = self.flat.device
        run("phc_adam_clip_step_dev", dev, self.flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.flat.numel(), self.lr_dev.data_ptr(), b1, b2, self.eps, self.step_dev.data_ptr(), self.max_norm,
            self.grad_norm.data_ptr(), self._ws.data_ptr(), self._ws.numel(), _stream(dev))
