from oracle.pyg import Data  # noqa: F401
