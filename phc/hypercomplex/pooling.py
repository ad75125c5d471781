from phc_gnn_b200.nn import PHMGlobalSumPooling, PHMSoftAttentionPooling  # noqa: F401
