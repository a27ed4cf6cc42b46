"""Import UNMODIFIED reference modules in the build container (test infrastructure only).

`/root/reference` cannot be imported as a package here: its __init__ pulls every model and therefore
every PyG layer.  We register empty parent packages whose __path__ points into the reference tree and
import single files on top of oracle/stubs (torch_geometric / dask stand-ins).  Nothing here exists on
the GPU box; callers must skip when REFERENCE_ROOT is absent.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "torch_geometric_temporal"))


def _ensure_pkg(name: str, path: str):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = [path]
        m.__package__ = name
        sys.modules[name] = m
    return sys.modules[name]


def load(modname: str):
    """e.g. load('nn.recurrent.dcrnn') -> module object of the reference file."""
    if not available():
        raise RuntimeError("reference tree not present")
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (_STUBS, repo_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    base = os.path.join(REFERENCE_ROOT, "torch_geometric_temporal")
    _ensure_pkg("torch_geometric_temporal", base)
    parts = modname.split(".")
    for i in range(1, len(parts)):
        _ensure_pkg("torch_geometric_temporal." + ".".join(parts[:i]), os.path.join(base, *parts[:i]))
    return importlib.import_module("torch_geometric_temporal." + modname)
