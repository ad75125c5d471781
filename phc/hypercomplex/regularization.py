from phc_gnn_b200.nn import phm_weight_regularization, multiplication_rule_regularization  # noqa: F401
