"""Minimal torch-geometric 1.6.1 surface used by the reference. Test infrastructure only."""
from . import typing, utils, data, nn, transforms  # noqa: F401
