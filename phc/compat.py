"""Opt-in compatibility shims for running the reference's UNCHANGED ``benchmarks/train_*.py`` on current PyTorch
(SURVEY.md D12).  Nothing here runs on ``import phc``; a launcher calls ``phc.compat.enable()`` (or sets ``PHC_COMPAT=1`` in the
environment of the script's process) before the script starts:

  * lr schedulers accept and ignore the removed ``verbose=`` keyword (train_hiv.py:287-289);
  * ``torch.load`` of the scripts' own whole-module pickles (train_hiv.py:369) needs ``weights_only=False``: ``trusted_load``
    is a context manager that switches the default for the files the script itself wrote, and only while it is active.
"""
import contextlib
import functools
import inspect
import os

import torch

_ENABLED = False


def _accept_verbose(cls):
    init = cls.__init__
    if "verbose" in inspect.signature(init).parameters or getattr(init, "_phc_compat", False):
        return

    @functools.wraps(init)
    def patched(self, *args, verbose=None, **kwargs):
        return init(self, *args, **kwargs)

    patched._phc_compat = True
    patched._phc_original = init
    cls.__init__ = patched


@contextlib.contextmanager
def trusted_load():
    """``with phc.compat.trusted_load(): model = torch.load(path)`` — full unpickling for a checkpoint the caller wrote itself."""
    original = torch.load

    @functools.wraps(original)
    def load(*args, **kwargs):
        kwargs.setdefault("weights_only", False)
        return original(*args, **kwargs)

    torch.load = load
    try:
        yield
    finally:
        torch.load = original


def enable(trust_checkpoints: bool = False) -> None:
    """Install the scheduler shim; ``trust_checkpoints=True`` additionally makes ``torch.load`` default to weights_only=False for
    the rest of the process (what the unchanged scripts need to reload their own ``model.pt``) — an explicit decision of the
    launcher, never a side effect of importing the package."""
    global _ENABLED
    for cls in (torch.optim.lr_scheduler.ReduceLROnPlateau, torch.optim.lr_scheduler.StepLR):
        _accept_verbose(cls)
    if trust_checkpoints:
        os.environ["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"
    _ENABLED = True


def enabled() -> bool:
    return _ENABLED
