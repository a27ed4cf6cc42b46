"""StaticGraphTemporalSignal -- drop-in for torch_geometric_temporal/signal/static_graph_temporal_signal.py
(:14-134): same constructor, `__getitem__(int|slice)`, iteration protocol, dtype rules and kwargs
passthrough.  B200 additions: the static graph tensors are wrapped ONCE (the reference re-wraps
edge_index / edge_weight on every snapshot, :62-72) and `device=` keeps them resident on the GPU, so
the plan cache of the layers hits on every snapshot."""
from typing import Sequence, Union

import numpy as np
import torch

from .data import Data

Edge_Index = Union[np.ndarray, None]
Edge_Weight = Union[np.ndarray, None]
Node_Features = Sequence[Union[np.ndarray, None]]
Targets = Sequence[Union[np.ndarray, None]]
Additional_Features = Sequence[np.ndarray]


def _typed(a):
    if a is None:
        return None
    if a.dtype.kind == "i":
        return torch.LongTensor(a)
    if a.dtype.kind == "f":
        return torch.FloatTensor(a)
    return None  # the reference falls through and returns None for other kinds (:85-88)


class StaticGraphTemporalSignal(object):
    def __init__(self, edge_index: Edge_Index, edge_weight: Edge_Weight, features: Node_Features, targets: Targets,
                 device=None, **kwargs: Additional_Features):
        self.edge_index = edge_index
        self.edge_weight = edge_weight
        self.features = features
        self.targets = targets
        self.device = device
        self.additional_feature_keys = []
        for key, value in kwargs.items():
            setattr(self, key, value)
            self.additional_feature_keys.append(key)
        self._check_temporal_consistency()
        self._set_snapshot_count()
        self._ei_t = None
        self._ew_t = None

    def _check_temporal_consistency(self):
        assert len(self.features) == len(self.targets), "Temporal dimension inconsistency."
        for key in self.additional_feature_keys:
            assert len(self.targets) == len(getattr(self, key)), "Temporal dimension inconsistency."

    def _set_snapshot_count(self):
        self.snapshot_count = len(self.features)

    def _put(self, t):
        return t if (t is None or self.device is None) else t.to(self.device, non_blocking=True)

    def _get_edge_index(self):
        if self.edge_index is None:
            return None
        if self._ei_t is None:
            self._ei_t = self._put(torch.LongTensor(self.edge_index))
        return self._ei_t

    def _get_edge_weight(self):
        if self.edge_weight is None:
            return None
        if self._ew_t is None:
            self._ew_t = self._put(torch.FloatTensor(self.edge_weight))
        return self._ew_t

    def _get_features(self, time_index: int):
        f = self.features[time_index]
        return None if f is None else self._put(torch.FloatTensor(f))

    def _get_target(self, time_index: int):
        return self._put(_typed(self.targets[time_index]))

    def _get_additional_features(self, time_index: int):
        return {k: self._put(_typed(getattr(self, k)[time_index])) for k in self.additional_feature_keys}

    def __getitem__(self, time_index: Union[int, slice]):
        if isinstance(time_index, slice):
            return StaticGraphTemporalSignal(
                self.edge_index, self.edge_weight, self.features[time_index], self.targets[time_index],
                device=self.device,
                **{key: getattr(self, key)[time_index] for key in self.additional_feature_keys})
        return Data(x=self._get_features(time_index), edge_index=self._get_edge_index(),
                    edge_attr=self._get_edge_weight(), y=self._get_target(time_index),
                    **self._get_additional_features(time_index))

    def __next__(self):
        if self.t < len(self.features):
            snapshot = self[self.t]
            self.t = self.t + 1
            return snapshot
        self.t = 0
        raise StopIteration

    def __iter__(self):
        self.t = 0
        return self

    def to(self, device):
        """Return a view of this signal whose snapshots are produced on `device`."""
        return StaticGraphTemporalSignal(self.edge_index, self.edge_weight, self.features, self.targets, device=device,
                                         **{k: getattr(self, k) for k in self.additional_feature_keys})
