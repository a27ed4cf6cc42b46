from oracle.pyg import glorot, zeros  # noqa: F401
