from oracle.pyg import ChebConv, GCNConv, MessagePassing  # noqa: F401
