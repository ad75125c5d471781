from phc_gnn_b200.nn import PHMSkipConnectAdd, PHMSkipConnectConcat  # noqa: F401
