"""Dynamic-graph snapshot iterators -- drop-ins for signal/dynamic_graph_temporal_signal.py (:13-138),
dynamic_graph_static_signal.py (:13-140) and their `...Batch` variants (dynamic_graph_temporal_signal_batch.py,
dynamic_graph_static_signal_batch.py) -- SURVEY 8f rank 4.

The four reference classes differ only in WHICH of the five fields (edge_index, edge_weight, x, y, batch) change over
time; they are expressed here as one table-driven iterator: a field is either `temporal` (a sequence indexed by the
snapshot) or `static` (one array shared by all snapshots, converted once).  Constructor signatures, public attribute
names (`edge_indices`, `edge_weights`, `features` / `feature`, `targets`, `batches`), `signal[t]`, `signal[a:b]`,
the iteration protocol (cursor reset on exhaustion) and the dtype rules (int -> LongTensor, float -> FloatTensor, None
passes through) are the reference's.

A changing `edge_index` means a fresh tensor per snapshot, so the layers' plan cache misses and the CSR operators are
rebuilt on the device for each snapshot (`stmp_plan_create`: ~30 small launches, no host round trip).  Snapshots of
one signal that reuse the same numpy array object for their edges (common: piecewise-constant graphs) share the
converted tensor, and therefore the plan."""
import collections
import zlib
from typing import Sequence, Union

import numpy as np
import torch

from .data import Data
from .static_graph_temporal_signal import _as_tensor

_FORCED = {"edge_index": torch.LongTensor, "edge_attr": torch.FloatTensor, "x": torch.FloatTensor, "batch": torch.LongTensor}


class _DynamicSignal(object):
    # (snapshot key, attribute name, temporal?) in constructor order; subclasses fill it in
    _FIELDS = ()

    def __init__(self, *values, device=None, **kwargs):
        if len(values) != len(self._FIELDS):
            raise TypeError(f"{type(self).__name__} takes {len(self._FIELDS)} positional arguments")
        for (_, attr, _), v in zip(self._FIELDS, values):
            setattr(self, attr, v)
        self.device = device
        self.additional_feature_keys = list(kwargs)
        for key, series in kwargs.items():
            setattr(self, key, series)
        lengths = {len(getattr(self, attr)) for _, attr, temporal in self._FIELDS if temporal}
        lengths |= {len(getattr(self, k)) for k in self.additional_feature_keys}
        assert len(lengths) == 1, "Temporal dimension inconsistency."
        self.snapshot_count = len(self.targets)
        self._converted = collections.OrderedDict()     # id(numpy array) -> (array, tensor): static fields and re-used per-snapshot arrays convert once
        self.t = 0

    def _place(self, tensor):
        if tensor is None or self.device is None:
            return tensor
        return tensor.to(self.device, non_blocking=True)

    def _tensor(self, key, array, share):
        if array is None:
            return None
        if share:
            # keyed on identity AND content: a caller may refill an edge buffer in place between snapshots (the reference
            # re-wraps the array on every __getitem__), so a stale tensor / plan must never be returned
            ck = (id(array), getattr(array, "shape", None), zlib.crc32(memoryview(np.ascontiguousarray(array)).cast("B")))
            hit = self._converted.get(ck)
            if hit is not None and hit[0] is array:
                self._converted.move_to_end(ck)
                return hit[1]
        t = self._place(_as_tensor(array, _FORCED.get(key)))
        if share:
            self._converted[ck] = (array, t)
            while len(self._converted) > 64:            # small LRU: bounds the device tensors a signal pins
                self._converted.popitem(last=False)
        return t

    def _snapshot(self, t: int) -> Data:
        fields = {}
        for key, attr, temporal in self._FIELDS:
            src = getattr(self, attr)
            # graph structure tensors are shared across snapshots that hand in the same array (plan cache hits)
            fields[key] = self._tensor(key, src[t] if temporal else src, share=(not temporal) or key in ("edge_index", "edge_attr", "batch"))
        extras = {k: self._place(_as_tensor(getattr(self, k)[t])) for k in self.additional_feature_keys}
        snap = Data(x=fields["x"], edge_index=fields["edge_index"], edge_attr=fields["edge_attr"], y=fields["y"], **extras)
        if "batch" in fields:
            snap.batch = fields["batch"]
            snap._keys.append("batch")
        return snap

    def _window(self, sl: slice):
        values = [getattr(self, attr)[sl] if temporal else getattr(self, attr) for _, attr, temporal in self._FIELDS]
        extras = {k: getattr(self, k)[sl] for k in self.additional_feature_keys}
        return type(self)(*values, device=self.device, **extras)

    def __getitem__(self, time_index: Union[int, slice]):
        return self._window(time_index) if isinstance(time_index, slice) else self._snapshot(time_index)

    def __len__(self):
        return self.snapshot_count

    def __iter__(self):
        self.t = 0
        return self

    def __next__(self):
        if self.t >= self.snapshot_count:
            self.t = 0
            raise StopIteration
        self.t += 1
        return self._snapshot(self.t - 1)

    def to(self, device):
        values = [getattr(self, attr) for _, attr, _ in self._FIELDS]
        extras = {k: getattr(self, k) for k in self.additional_feature_keys}
        return type(self)(*values, device=device, **extras)


class DynamicGraphTemporalSignal(_DynamicSignal):
    """(edge_indices, edge_weights, features, targets, **kwargs): graph, features and targets all change per snapshot."""
    _FIELDS = (("edge_index", "edge_indices", True), ("edge_attr", "edge_weights", True), ("x", "features", True), ("y", "targets", True))


class DynamicGraphStaticSignal(_DynamicSignal):
    """(edge_indices, edge_weights, feature, targets, **kwargs): the graph changes, the node features are one array."""
    _FIELDS = (("edge_index", "edge_indices", True), ("edge_attr", "edge_weights", True), ("x", "feature", False), ("y", "targets", True))


class DynamicGraphTemporalSignalBatch(_DynamicSignal):
    """(edge_indices, edge_weights, features, targets, batches, **kwargs): as above with a per-snapshot `batch` vector."""
    _FIELDS = DynamicGraphTemporalSignal._FIELDS + (("batch", "batches", True),)


class DynamicGraphStaticSignalBatch(_DynamicSignal):
    """(edge_indices, edge_weights, feature, targets, batches, **kwargs)."""
    _FIELDS = DynamicGraphStaticSignal._FIELDS + (("batch", "batches", True),)
