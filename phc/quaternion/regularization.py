from phc_gnn_b200.nn import get_model_blocks  # noqa: F401


def quaternion_weight_regularization(model, device=None, p: int = 1):
    raise NotImplementedError("the quaternion model family is out of scope of the B200 hot path (SURVEY.md §2)")
