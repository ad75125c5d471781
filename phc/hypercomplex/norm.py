from phc_gnn_b200.nn import NaivePHMNorm, PHMNorm  # noqa: F401
