"""reference phc/quaternion/regularization.py."""
from phc_gnn_b200.nn import get_model_blocks  # noqa: F401
from phc_gnn_b200.quaternion import quaternion_weight_regularization  # noqa: F401
