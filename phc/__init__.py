"""Drop-in ``phc`` namespace: the reference's import paths (``phc.hypercomplex.*``, and the few
``phc.quaternion`` helpers its training scripts import) re-exported from ``phc_gnn_b200``.

Importing it also installs two small compatibility shims so that the reference's unchanged
``benchmarks/train_*.py`` run on current PyTorch:
  * lr schedulers accept and ignore the removed ``verbose=`` keyword (train_hiv.py:287-289);
  * ``torch.load`` of whole pickled modules (train_hiv.py:369) defaults to weights_only=False.
"""
import functools
import inspect
import os

import torch


def _accept_verbose(cls):
    init = cls.__init__
    if "verbose" in inspect.signature(init).parameters or getattr(init, "_phc_compat", False):
        return

    @functools.wraps(init)
    def patched(self, *args, verbose=None, **kwargs):
        return init(self, *args, **kwargs)

    patched._phc_compat = True
    cls.__init__ = patched


for _cls in (torch.optim.lr_scheduler.ReduceLROnPlateau, torch.optim.lr_scheduler.StepLR):
    _accept_verbose(_cls)
os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
