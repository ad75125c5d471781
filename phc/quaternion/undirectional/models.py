"""reference phc/quaternion/undirectional/models.py — the quaternion family on the PHM kernels (n = 4, fixed Hamilton rule)."""
from phc_gnn_b200.quaternion import QuaternionSkipConnectAdd, QuaternionSkipConnectConcat  # noqa: F401
