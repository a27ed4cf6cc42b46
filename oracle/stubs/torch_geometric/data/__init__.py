from oracle.pyg import Data  # noqa: F401

Batch = Data      # attribute bag with a `batch` keyword: all the reference iterators need (static/dynamic ...Batch signals)
