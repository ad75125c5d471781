from phc_gnn_b200.nn import IntegerEncoder  # noqa: F401
