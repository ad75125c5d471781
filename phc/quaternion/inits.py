"""reference phc/quaternion/inits.py ([4, in, out] stacks instead of four [out, in] tensors for the quaternion initialisers)."""
from phc_gnn_b200.functional import glorot_normal, glorot_uniform  # noqa: F401
from phc_gnn_b200.quaternion import quaternion_init, quaternion_orthogonal_init as orthogonal_init  # noqa: F401
