"""StaticGraphTemporalSignalBatch -- drop-in for signal/static_graph_temporal_signal_batch.py (:14-146; SURVEY 8f
rank 2): a static graph that is itself a mini-batch of graphs (block-diagonal `edge_index` + a `batches` vector
mapping nodes to graphs).  Snapshots carry `.batch` next to `x / edge_index / edge_attr / y`.

Built on StaticGraphTemporalSignal: the three static tensors (edge_index, edge_weight, batch) are converted once and
shared by every snapshot, so layers keyed on tensor identity reuse one cached plan for the whole sequence."""
from typing import Union

import numpy as np
import torch

from .static_graph_temporal_signal import StaticGraphTemporalSignal, _as_tensor, Edge_Index, Edge_Weight, Node_Features, Targets

Batches = Union[np.ndarray, None]


class StaticGraphTemporalSignalBatch(StaticGraphTemporalSignal):
    def __init__(self, edge_index: Edge_Index, edge_weight: Edge_Weight, features: Node_Features, targets: Targets,
                 batches: Batches, device=None, **kwargs):
        super().__init__(edge_index, edge_weight, features, targets, device=device, **kwargs)
        self.batches = batches
        self._batch_tensor = None

    def _batch_index(self):
        if self.batches is not None and self._batch_tensor is None:
            self._batch_tensor = self._place(_as_tensor(self.batches, torch.LongTensor))
        return self._batch_tensor

    def _snapshot(self, t: int):
        snap = super()._snapshot(t)
        snap.batch = self._batch_index()
        snap._keys.append("batch")          # so Data.to(device) carries it
        return snap

    def _window(self, sl: slice) -> "StaticGraphTemporalSignalBatch":
        extras = {k: getattr(self, k)[sl] for k in self.additional_feature_keys}
        return StaticGraphTemporalSignalBatch(self.edge_index, self.edge_weight, self.features[sl], self.targets[sl],
                                              self.batches, device=self.device, **extras)

    def to(self, device) -> "StaticGraphTemporalSignalBatch":
        extras = {k: getattr(self, k) for k in self.additional_feature_keys}
        return StaticGraphTemporalSignalBatch(self.edge_index, self.edge_weight, self.features, self.targets, self.batches,
                                              device=device, **extras)
