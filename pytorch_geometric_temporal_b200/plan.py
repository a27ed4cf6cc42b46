"""GraphPlan: Python owner of an `stmp_plan` (cached, normalised graph operators on the device).

Replaces the per-call renormalisation in the reference (dcrnn.py:59-77 dense adjacency + nonzero,
PyG get_laplacian/gcn_norm on every ChebConv/GCNConv call) and BatchedDCRNN's `torch.equal`
freshness test (dcrnn.py:446-447, a device sync per forward): freshness is decided on the host from
(data_ptr, _version, shape) of edge_index / edge_weight.
"""
import ctypes
from typing import Optional

import torch

from . import _lib


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: pytorch_geometric_temporal_b200 runs the hot path on CUDA only "
            "(hand-written sm_100a kernels, no CPU fallback).")


class GraphPlan:
    def __init__(self, flavor: int, edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int,
                 normalization=None, lambda_max: Optional[float] = None, flags: int = 0,
                 lambda_node: Optional[torch.Tensor] = None):
        _require_cuda(edge_index, "edge_index")
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError(f"edge_index must have shape [2, E], got {tuple(edge_index.shape)}")
        if normalization not in _lib.NORM_CODE:
            raise AssertionError("Invalid normalization")  # astgcn.py:62
        ei = edge_index.to(torch.int64).contiguous()
        ew = None
        if edge_weight is not None:
            _require_cuda(edge_weight, "edge_weight")
            ew = edge_weight.detach().to(torch.float32).contiguous()
            if ew.numel() != ei.size(1):
                raise RuntimeError(f"edge_weight has {ew.numel()} entries for {ei.size(1)} edges")
        self.flavor, self.num_nodes, self.num_edges = flavor, int(num_nodes), int(ei.size(1))
        self.device = ei.device
        self._h = ctypes.c_void_p()
        lam = -1.0 if lambda_max is None else float(lambda_max)
        with torch.cuda.device(ei.device):
            if lambda_node is not None:      # per-graph lambda_max of a multi-graph mini-batch, already expanded to nodes
                ln = lambda_node.detach().to(device=ei.device, dtype=torch.float32).contiguous()
                if ln.numel() != self.num_nodes:
                    raise RuntimeError(f"lambda_max[batch] has {ln.numel()} entries for {self.num_nodes} nodes")
                rc = _lib.lib().stmp_plan_create_pergraph(flavor, self.num_nodes, self.num_edges, _lib.ptr(ei), _lib.ptr(ew),
                                                          _lib.NORM_CODE[normalization], _lib.ptr(ln), flags, _lib.stream_ptr(),
                                                          ctypes.byref(self._h))
            else:
                rc = _lib.lib().stmp_plan_create(flavor, self.num_nodes, self.num_edges, _lib.ptr(ei), _lib.ptr(ew),
                                                 _lib.NORM_CODE[normalization], lam, flags, _lib.stream_ptr(),
                                                 ctypes.byref(self._h))
        _lib.check(rc)
        self.n_ops = _lib.lib().stmp_plan_num_ops(self._h)

    @property
    def handle(self):
        return self._h

    def nnz(self, op: int = 0) -> int:
        return int(_lib.lib().stmp_plan_nnz(self._h, op))

    def export(self, op: int = 0, transposed: bool = False):
        """(rowptr, col, val, eid) as torch tensors -- test/introspection helper."""
        n, nnz = self.num_nodes, self.nnz(op)
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=self.device)
        col = torch.empty(nnz, dtype=torch.int32, device=self.device)
        val = torch.empty(nnz, dtype=torch.float32, device=self.device)
        eid = torch.empty(nnz, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().stmp_plan_export(self._h, op, int(transposed), _lib.ptr(rowptr), _lib.ptr(col),
                                                   _lib.ptr(val), _lib.ptr(eid), _lib.stream_ptr()))
        return rowptr, col, val, eid

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().stmp_plan_destroy(h)
            except Exception:
                pass
            self._h = None


class PlanCache:
    """Per-module cache keyed on the identity/version of the graph tensors (no device sync)."""

    def __init__(self, max_entries: int = 4):
        self._entries = {}
        self._max = max_entries

    @staticmethod
    def _tkey(t):
        return None if t is None else (t.data_ptr(), t._version, tuple(t.shape), t.dtype, t.device)

    def get(self, flavor, edge_index, edge_weight, num_nodes, normalization=None, lambda_max=None, flags=0, batch=None) -> GraphPlan:
        """`lambda_max`: None, a host scalar / 0-d tensor, or -- with the node->graph vector `batch` -- one value per graph
        (PyG: `lambda_max[batch[edge_index[0]]]`)."""
        lam, lam_node, extra = None, None, ()
        if lambda_max is not None:
            if torch.is_tensor(lambda_max) and lambda_max.numel() > 1:
                if batch is None:
                    raise ValueError("a lambda_max vector needs the `batch` vector of the mini-batch (one graph id per node)")
                lam_node = lambda_max.to(torch.float32)[batch]
                extra = (self._tkey(lambda_max), self._tkey(batch))
            else:
                lam = float(lambda_max)  # scalar lambda_max (host value or 0-d tensor; a sync only if it is a tensor)
        key = (flavor, self._tkey(edge_index), self._tkey(edge_weight), int(num_nodes), normalization, lam, flags) + extra
        hit = self._entries.get(key)
        if hit is not None:
            return hit[0]
        plan = GraphPlan(flavor, edge_index, edge_weight, num_nodes, normalization, lam, flags, lambda_node=lam_node)
        if len(self._entries) >= self._max:
            self._entries.pop(next(iter(self._entries)))
        # keep the keyed tensors alive so their addresses cannot be recycled while the entry exists
        self._entries[key] = (plan, edge_index, edge_weight, lambda_max if extra else None, batch if extra else None)
        return plan
