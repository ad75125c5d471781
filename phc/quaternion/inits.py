from phc_gnn_b200.functional import glorot_normal, glorot_uniform  # noqa: F401
