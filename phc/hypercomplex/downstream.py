from phc_gnn_b200.nn import PHMDownstreamNet  # noqa: F401
