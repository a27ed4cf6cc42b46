from oracle.pyg import (to_dense_adj, dense_to_sparse, remove_self_loops, add_self_loops,  # noqa: F401
                        add_remaining_self_loops, get_laplacian)
