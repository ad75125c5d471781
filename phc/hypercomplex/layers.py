from phc_gnn_b200.nn import PHMLinear, PHMMLP, RealTransformer, phm_dropout  # noqa: F401
