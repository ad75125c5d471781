from phc_gnn_b200.functional import phm_init, unitary_init  # noqa: F401
