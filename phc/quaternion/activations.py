from phc_gnn_b200.nn import get_module_activation  # noqa: F401
