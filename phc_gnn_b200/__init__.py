"""phc_gnn_b200 — B200-native (sm_100a) implementation of the PHC-GNN hypercomplex message-passing
hot path behind the reference's ``phc.hypercomplex`` module API.

(The task names the package ``phc-gnn_b200``; a hyphen is not importable, hence the underscore.)
"""
from . import _lib  # noqa: F401

__all__ = ["build", "load_library"]


def build(force: bool = False, verbose: bool = False) -> str:
    return _lib.build(force=force, verbose=verbose)


def load_library():
    return _lib.load()
