"""Minimal snapshot container standing in for torch_geometric.data.Data (the reference uses it purely
as an attribute bag: signal/static_graph_temporal_signal.py:119-120; tests read .x/.edge_index/
.edge_attr/.y).  Adds `.to(device)` / `.cuda()` so a snapshot can be moved in one call."""
import torch

# Static-graph tensors (edge_index / edge_attr) are the same host objects on every snapshot of a
# StaticGraphTemporalSignal; moving each snapshot with `.to(device)` must not mint a new device copy per
# snapshot (the layers key their cached plans on tensor identity).  Small memo: host tensor -> device copy.
_GRAPH_KEYS = ("edge_index", "edge_attr")
_DEVICE_MEMO = {}


def _memo_to(t, device, non_blocking):
    key = (id(t), t._version, str(device))
    hit = _DEVICE_MEMO.get(key)
    if hit is not None and hit[0] is t:
        return hit[1]
    if len(_DEVICE_MEMO) >= 16:
        _DEVICE_MEMO.pop(next(iter(_DEVICE_MEMO)))
    d = t.to(device, non_blocking=non_blocking)
    _DEVICE_MEMO[key] = (t, d)
    return d


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        self._keys = ["x", "edge_index", "edge_attr", "y"]
        for k, v in kwargs.items():
            setattr(self, k, v)
            self._keys.append(k)

    def keys(self):
        return [k for k in self._keys if getattr(self, k) is not None]

    def to(self, device, non_blocking=False):
        out = Data()
        out._keys = list(self._keys)
        for k in self._keys:
            v = getattr(self, k)
            if torch.is_tensor(v):
                v = _memo_to(v, device, non_blocking) if k in _GRAPH_KEYS else v.to(device, non_blocking=non_blocking)
            setattr(out, k, v)
        return out

    def cuda(self, device=None, non_blocking=False):
        return self.to(torch.device("cuda" if device is None else device), non_blocking)

    @property
    def num_nodes(self):
        if self.x is not None:
            return self.x.size(0)
        return int(self.edge_index.max()) + 1

    def __repr__(self):
        parts = [f"{k}={list(getattr(self, k).shape) if torch.is_tensor(getattr(self, k)) else getattr(self, k)}" for k in self.keys()]
        return f"Data({', '.join(parts)})"
