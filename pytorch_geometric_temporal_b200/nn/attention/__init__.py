from .astgcn import ASTGCN, ASTGCNBlock, ChebConvAttention, SpatialAttention, TemporalAttention  # noqa: F401
