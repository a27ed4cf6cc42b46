"""Index-batching data feed -- drop-in for signal/index_dataset.py (:8-57) plus the loader arithmetic of
dataset/metr_la.py:194-234, re-designed for a GPU-resident series.

* `IndexDataset(indices, data, horizon, lazy=False, gpu=False)`: same constructor / `__getitem__`
  (x = data[i:i+h], y = data[i+h:i+2h]); `dask` is not required (lazy=True is rejected).
* `IndexBatchLoader`: replaces `DataLoader(IndexDataset)` + default collate + per-batch H2D.  The whole
  series stays in HBM; one launch of `stmp_window_gather` materialises a (B,h,N,F) batch, or -- for the
  fused DCRNN kernel -- nothing is materialised and only the window starts are handed to the kernel.
  Rank sharding follows torch DistributedSampler semantics (metr_la.py:220-228): seeded permutation per
  epoch, padded to a multiple of world size, rank r takes perm[r::world].
"""
import math

import numpy as np
import torch

from .. import ops


class IndexDataset(torch.utils.data.Dataset):
    def __init__(self, indices, data, horizon, lazy=False, gpu=False):
        if lazy:
            raise ValueError("lazy (Dask) index-batching is out of scope: the series is kept resident in HBM instead")
        self.indices = indices
        self.data = data
        self.horizon = horizon
        self.lazy = lazy
        self.gpu = gpu

    def __len__(self):
        return self.indices.shape[0]

    def __getitem__(self, x):
        idx = self.indices[x]
        y_start = idx + self.horizon
        if self.gpu:
            return self.data[idx:y_start, ...], self.data[y_start:y_start + self.horizon, ...]
        return torch.from_numpy(self.data[idx:y_start, ...]), torch.from_numpy(self.data[y_start:y_start + self.horizon, ...])


def shard_indices(n: int, world_size: int, rank: int, shuffle: bool, seed: int, epoch: int, drop_last: bool = False):
    """torch.utils.data.DistributedSampler index selection (host, bit-exact): returns positions in [0,n)."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(n, generator=g).tolist()
    else:
        order = list(range(n))
    if world_size <= 1:
        return order
    if drop_last and n % world_size != 0:
        num = math.ceil((n - world_size) / world_size)
    else:
        num = math.ceil(n / world_size)
    total = num * world_size
    if not drop_last:
        pad = total - len(order)
        if pad <= len(order):
            order += order[:pad]
        else:
            order += (order * math.ceil(pad / len(order)))[:pad]
    else:
        order = order[:total]
    return order[rank:total:world_size]


class IndexBatchLoader(object):
    """Iterates batches of windows over a device-resident series.

    Args:
        series: (T_total, N, F) float tensor on the GPU (already normalised).
        indices: 1-D array of window start positions (the reference's `x_train` etc.).
        horizon: window length h (x = h steps, y = next h steps).
        batch_size, shuffle, world_size, rank, seed: as DataLoader / DistributedSampler.
        materialize: True -> yields (x, y) tensors (B,h,N,F) gathered by `stmp_window_gather`;
                     False -> yields (start, None): int64 device tensor of window starts for kernels
                     that read the series in place (BatchedDCRNN.forward_indexed).
    """

    def __init__(self, series, indices, horizon, batch_size, shuffle=False, world_size=1, rank=0, seed=0,
                 drop_last=False, materialize=True):
        if not series.is_cuda:
            raise RuntimeError("IndexBatchLoader keeps the series resident on the GPU; move it with .cuda() first")
        self.series = series.contiguous()
        self.indices = np.asarray(indices, dtype=np.int64)
        self.horizon, self.batch_size, self.shuffle = int(horizon), int(batch_size), shuffle
        self.world_size, self.rank, self.seed, self.drop_last = world_size, rank, seed, drop_last
        self.materialize = materialize
        self.epoch = 0

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def _positions(self):
        return shard_indices(len(self.indices), self.world_size, self.rank, self.shuffle, self.seed, self.epoch)

    def __len__(self):
        if self.world_size > 1:
            n = math.ceil(len(self.indices) / self.world_size)     # DistributedSampler pads every shard to this length
        else:
            n = len(self.indices)
        return n // self.batch_size if self.drop_last else math.ceil(n / self.batch_size)

    def host_batches(self):
        """Host-side iteration: pinned int64 window-start tensors, one per batch (what a `DevicePrefetcher` ships to the
        GPU -- B x 8 bytes per step instead of the B x h x N x F windows themselves)."""
        pos = self._positions()
        starts = torch.from_numpy(self.indices[pos])
        if torch.cuda.is_available():
            starts = starts.pin_memory()
        for s in range(0, starts.numel(), self.batch_size):
            st = starts[s:s + self.batch_size]
            if self.drop_last and st.numel() < self.batch_size:
                break
            yield st

    def __iter__(self):
        pos = self._positions()
        starts = torch.from_numpy(self.indices[pos]).pin_memory() if torch.cuda.is_available() else torch.from_numpy(self.indices[pos])
        starts = starts.to(self.series.device, non_blocking=True)
        n = starts.numel()
        for s in range(0, n, self.batch_size):
            st = starts[s:s + self.batch_size]
            if self.drop_last and st.numel() < self.batch_size:
                break
            if self.materialize:
                yield ops.window_gather(self.series, st, self.horizon, with_target=True)
            else:
                yield st, None


def index_splits(t_total: int, lags: int, ratio=(0.7, 0.1, 0.2)):
    """Window-start arrays as dataset/metr_la.py:204-213 builds them."""
    x_i = np.arange(t_total - (2 * lags - 1))
    n = x_i.shape[0]
    n_tr, n_te = round(n * ratio[0]), round(n * ratio[2])
    n_va = n - n_tr - n_te
    return x_i[:n_tr], x_i[n_tr:n_tr + n_va], x_i[-n_te:]


class DevicePrefetcher(object):
    """Wraps an iterator of pinned host batches (tensors or tuples of tensors): the H2D copy of batch i+1 is
    issued on a side stream while batch i is being consumed on the current stream.  Replaces the per-batch
    blocking `.to(device)` of the reference's training loops (examples/indexBatching/DCRNN/pems_ddp.py:104-108)."""

    def __init__(self, it, device):
        self.it, self.device = iter(it), torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs, self.slot = [None, None], 0   # two staging buffers per tensor position
        self.next = None
        self._preload()

    def _staging(self, t, pos):
        """Staging buffer of the current slot for tensor position `pos` (allocated on the CONSUMER's stream, so the
        caching allocator accounts it to the stream that reads it)."""
        key = (pos, tuple(t.shape), t.dtype)
        bank = self.bufs[self.slot]
        if bank is None:
            bank = self.bufs[self.slot] = {}
        buf = bank.get(key)
        if buf is None:
            buf = bank[key] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        return buf

    def _preload(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self.next = None
            return
        tensors = (batch,) if torch.is_tensor(batch) else tuple(batch)
        cur = torch.cuda.current_stream(self.device)          # the consumer's stream -- taken BEFORE switching streams
        bufs = tuple(self._staging(t, i) for i, t in enumerate(tensors))
        # the staging buffers of this slot were last read two batches ago by kernels already enqueued on the consumer's
        # stream: the copy that overwrites them must wait for everything enqueued there so far
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            for buf, t in zip(bufs, tensors):
                buf.copy_(t, non_blocking=True)
        self.next = bufs[0] if torch.is_tensor(batch) else bufs
        self.slot ^= 1

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        torch.cuda.current_stream(self.device).wait_stream(self.copy_stream)
        batch = self.next
        self._preload()
        return batch
