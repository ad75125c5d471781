"""Drop-in ``phc`` namespace: the reference's import paths (``phc.hypercomplex.*``, and the few
``phc.quaternion`` helpers its training scripts import) re-exported from ``phc_gnn_b200``.

Importing it changes nothing in torch.  The shims the reference's unchanged ``benchmarks/train_*.py`` need on current
PyTorch live in ``phc.compat`` and are opt-in: ``phc.compat.enable()``, or ``PHC_COMPAT=1`` /
``PHC_COMPAT=trust-checkpoints`` in the environment of the process that runs the script.
"""
import os

_mode = os.environ.get("PHC_COMPAT", "")
if _mode:
    from . import compat as _compat
    _compat.enable(trust_checkpoints=_mode == "trust-checkpoints")
