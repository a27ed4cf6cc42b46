"""temporal_signal_split -- drop-in for signal/train_test_split.py:36-54."""


def temporal_signal_split(data_iterator, train_ratio: float = 0.8):
    train_snapshots = int(train_ratio * data_iterator.snapshot_count)
    return data_iterator[0:train_snapshots], data_iterator[train_snapshots:]
