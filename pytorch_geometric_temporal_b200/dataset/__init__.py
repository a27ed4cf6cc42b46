from . import synthetic  # noqa: F401
from .chickenpox import ChickenpoxDatasetLoader  # noqa: F401
