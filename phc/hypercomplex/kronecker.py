from phc_gnn_b200.functional import (kronecker_product, kronecker_product_einsum_batched,  # noqa: F401
                                     kronecker_product_single)
