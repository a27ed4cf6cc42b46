from .dcrnn import DConv, DCRNN, BatchedDConv, BatchedDCRNN  # noqa: F401
