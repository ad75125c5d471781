from phc_gnn_b200.nn import (PHMConv, PHMGINEConv, PHMConvSoftmax, PHMGINEConvSoftmax,  # noqa: F401
                             PHMPNAConvSimple, PHMMessagePassing)
