from phc_gnn_b200.nn import PHMEncoder, NaivePHMEncoder  # noqa: F401
