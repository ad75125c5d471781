from phc_gnn_b200.functional import get_multiplication_matrices, phm_cat  # noqa: F401
