"""ChickenpoxDatasetLoader -- offline drop-in for dataset/chickenpox.py (:13-132).  The reference
downloads chickenpox.json (:32-38); this loader reads the same content from the .npz shipped inside the
package (dataset/data/chickenpox.npz: the dataset's `edges` and `FX` arrays re-encoded by
tests/golden/make_goldens.py) or from a user-supplied .npz/.json.  20 nodes, 102 edges (weights 1), FX (521,20)."""
import json
import os

import numpy as np
import torch

from ..signal import StaticGraphTemporalSignal, IndexDataset

_DEFAULT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "chickenpox.npz")


class ChickenpoxDatasetLoader(object):
    def __init__(self, index: bool = False, path: str = None):
        path = path or os.environ.get("STMP_CHICKENPOX", _DEFAULT)
        if path.endswith(".json"):
            with open(path) as f:
                d = json.load(f)
            self._dataset = {"edges": np.array(d["edges"], dtype=np.int64), "FX": np.array(d["FX"])}
        else:
            z = np.load(path)
            self._dataset = {"edges": z["edges"], "FX": z["FX"]}
        self.index = index

    def _get_edges(self):
        self._edges = np.array(self._dataset["edges"]).T

    def _get_edge_weights(self):
        self._edge_weights = np.ones(self._edges.shape[1])

    def _get_targets_and_features(self):
        st = np.array(self._dataset["FX"])
        self.features = [st[i:i + self.lags, :].T for i in range(st.shape[0] - self.lags)]
        self.targets = [st[i + self.lags, :].T for i in range(st.shape[0] - self.lags)]

    def get_dataset(self, lags: int = 4, device=None) -> StaticGraphTemporalSignal:
        self.lags = lags
        self._get_edges()
        self._get_edge_weights()
        self._get_targets_and_features()
        return StaticGraphTemporalSignal(self._edges, self._edge_weights, self.features, self.targets, device=device)

    def get_index_dataset(self, lags=4, batch_size=4, shuffle=False, allGPU=-1, ratio=(0.7, 0.1, 0.2), dask_batching=False):
        """Same return tuple as chickenpox.py:74-132 (three DataLoaders over IndexDataset, edges, weights)."""
        if not self.index:
            raise ValueError("get_index_dataset requires 'index=True' in the constructor.")
        data = np.array(self._dataset["FX"])
        edges = torch.tensor(self._dataset["edges"], dtype=torch.int64).T
        edge_weights = torch.ones(edges.shape[1], dtype=torch.float)
        num_samples = data.shape[0]
        if allGPU != -1:
            data = torch.tensor(data, dtype=torch.float).to(f"cuda:{allGPU}").unsqueeze(-1)
        else:
            data = np.expand_dims(data, axis=-1)
        x_i = np.arange(num_samples - (2 * lags - 1))
        n = x_i.shape[0]
        n_tr, n_te = round(n * ratio[0]), round(n * ratio[2])
        n_va = n - n_tr - n_te
        mk = lambda idx: torch.utils.data.DataLoader(IndexDataset(idx, data, lags, gpu=not (allGPU == -1), lazy=dask_batching),
                                                     batch_size=batch_size, shuffle=shuffle)
        return mk(x_i[:n_tr]), mk(x_i[n_tr:n_tr + n_va]), mk(x_i[-n_te:]), edges, edge_weights
