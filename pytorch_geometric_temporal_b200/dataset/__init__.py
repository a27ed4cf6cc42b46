from . import synthetic  # noqa: F401
from .chickenpox import ChickenpoxDatasetLoader  # noqa: F401
from .traffic import METRLADatasetLoader, PemsBayDatasetLoader, dense_to_sparse  # noqa: F401
