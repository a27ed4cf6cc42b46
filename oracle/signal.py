"""CPU restatement of the data feed: signal/static_graph_temporal_signal.py, signal/index_dataset.py,
signal/train_test_split.py, and the index arithmetic of dataset/metr_la.py:194-213 (oracle)."""
import numpy as np
import torch


def snapshot(edge_index, edge_weight, features, targets, t):
    """StaticGraphTemporalSignal.__getitem__(int) (static_graph_temporal_signal.py:62-121) as a dict."""
    def tgt(a):
        if a is None:
            return None
        return torch.LongTensor(a) if a.dtype.kind == "i" else torch.FloatTensor(a)
    return dict(
        x=None if features[t] is None else torch.FloatTensor(features[t]),
        edge_index=None if edge_index is None else torch.LongTensor(edge_index),
        edge_attr=None if edge_weight is None else torch.FloatTensor(edge_weight),
        y=tgt(targets[t]),
    )


def index_window(data, indices, i, horizon):
    """IndexDataset.__getitem__ (index_dataset.py:32-57)."""
    s = int(indices[i])
    return data[s:s + horizon, ...], data[s + horizon:s + 2 * horizon, ...]


def split_counts(snapshot_count, ratio):
    """temporal_signal_split (train_test_split.py:49-52)."""
    n = int(ratio * snapshot_count)
    return n, snapshot_count - n


def index_splits(t_total, lags, ratio=(0.7, 0.1, 0.2)):
    """dataset/metr_la.py:204-213: indices=arange(T-(2*lags-1)); train/val/test index arrays."""
    x_i = np.arange(t_total - (2 * lags - 1))
    n = x_i.shape[0]
    n_tr, n_te = round(n * ratio[0]), round(n * ratio[2])
    n_va = n - n_tr - n_te
    return x_i[:n_tr], x_i[n_tr:n_tr + n_va], x_i[-n_te:]


def zscore(data, axes):
    """dataset/metr_la.py:194-198 (index path): mean/std over (time, nodes) per feature."""
    means = np.mean(data, axis=axes)
    stds = np.std(data, axis=axes)
    return (data - means) / stds, means, stds
