"""StaticGraphTemporalSignal -- the snapshot iterator of the reference
(torch_geometric_temporal/signal/static_graph_temporal_signal.py:14-134) re-built around a device-resident
static graph.

Contract kept (constructor, public attributes, `signal[t]`, `signal[a:b]`, iteration protocol, dtype rules of
:62-100, extra per-snapshot arrays passed as keyword arguments).  What differs underneath:

* the graph is static, so `edge_index` / `edge_weight` are converted ONCE and every snapshot hands out the same
  tensor objects (the reference re-wraps both arrays on every `__getitem__`, :62-72) -- the layers key their cached
  plans on tensor identity, so this is what makes the plan cache hit on every snapshot;
* `device=` (or `.to(device)`) keeps those tensors, and the snapshots it produces, on the GPU.
"""
from typing import Optional, Sequence, Union

import numpy as np
import torch

from .data import Data

Edge_Index = Union[np.ndarray, None]
Edge_Weight = Union[np.ndarray, None]
Node_Features = Sequence[Union[np.ndarray, None]]
Targets = Sequence[Union[np.ndarray, None]]
Additional_Features = Sequence[np.ndarray]

# numpy kind -> torch constructor; anything else (bool, complex, ...) yields None, as in the reference (:85-100)
_BY_KIND = {"i": torch.LongTensor, "f": torch.FloatTensor}


def _as_tensor(array: Optional[np.ndarray], force=None):
    if array is None:
        return None
    ctor = force or _BY_KIND.get(array.dtype.kind)
    return None if ctor is None else ctor(array)


class StaticGraphTemporalSignal(object):
    def __init__(self, edge_index: Edge_Index, edge_weight: Edge_Weight, features: Node_Features, targets: Targets,
                 device=None, **kwargs: Additional_Features):
        self.edge_index, self.edge_weight = edge_index, edge_weight
        self.features, self.targets = features, targets
        self.device = device
        self.additional_feature_keys = list(kwargs)
        for key, series in kwargs.items():
            setattr(self, key, series)
        lengths = {len(self.features), len(self.targets), *(len(getattr(self, k)) for k in self.additional_feature_keys)}
        assert len(lengths) == 1, "Temporal dimension inconsistency."
        self.snapshot_count = len(self.features)
        self._graph = None  # (edge_index tensor, edge_weight tensor), built lazily, shared by all snapshots
        self.t = 0

    # ---- static part -------------------------------------------------------------------------------------------
    def _place(self, tensor):
        if tensor is None or self.device is None:
            return tensor
        return tensor.to(self.device, non_blocking=True)

    def _static_graph(self):
        if self._graph is None:
            self._graph = (self._place(_as_tensor(self.edge_index, torch.LongTensor)),
                           self._place(_as_tensor(self.edge_weight, torch.FloatTensor)))
        return self._graph

    # ---- access -------------------------------------------------------------------------------------------------
    def _snapshot(self, t: int) -> Data:
        ei, ew = self._static_graph()
        extras = {k: self._place(_as_tensor(getattr(self, k)[t])) for k in self.additional_feature_keys}
        return Data(x=self._place(_as_tensor(self.features[t], torch.FloatTensor)), edge_index=ei, edge_attr=ew,
                    y=self._place(_as_tensor(self.targets[t])), **extras)

    def _window(self, sl: slice) -> "StaticGraphTemporalSignal":
        extras = {k: getattr(self, k)[sl] for k in self.additional_feature_keys}
        return StaticGraphTemporalSignal(self.edge_index, self.edge_weight, self.features[sl], self.targets[sl],
                                         device=self.device, **extras)

    def __getitem__(self, time_index: Union[int, slice]):
        return self._window(time_index) if isinstance(time_index, slice) else self._snapshot(time_index)

    def __len__(self):
        return self.snapshot_count

    def __iter__(self):
        self.t = 0
        return self

    def __next__(self):
        if self.t >= self.snapshot_count:
            self.t = 0          # the reference resets its cursor when exhausted (:128-134): a second epoch just works
            raise StopIteration
        self.t += 1
        return self._snapshot(self.t - 1)

    def to(self, device) -> "StaticGraphTemporalSignal":
        """A view of this signal whose snapshots are produced on `device`."""
        extras = {k: getattr(self, k) for k in self.additional_feature_keys}
        return StaticGraphTemporalSignal(self.edge_index, self.edge_weight, self.features, self.targets, device=device, **extras)
