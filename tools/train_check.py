#!/usr/bin/env python
"""Trajectory check of the training probe's ingredients: the same 12 optimisation steps with {fused, reference-form}
masked-MAE loss x {fused, foreach} Adam, eager.  Prints the loss sequence of every combination (they must agree to
rounding) -- guards against a fused component silently changing the optimisation."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_geometric_temporal_b200 import distributed as D  # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic  # noqa: E402
from pytorch_geometric_temporal_b200.nn.recurrent import BatchedDCRNN  # noqa: E402
from pytorch_geometric_temporal_b200.signal import IndexBatchLoader, index_splits  # noqa: E402


def run(fused_loss, fused_adam, steps=12):
    dev = torch.device("cuda")
    ei, ew, series = synthetic.metr_la_like(0, 2048)
    ei, ew, series = torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev), torch.from_numpy(series).to(dev)
    torch.manual_seed(0)
    model, head = BatchedDCRNN(2, 32, 2).to(dev), torch.nn.Linear(32, 1).to(dev)
    params = list(model.parameters()) + list(head.parameters())
    sync = D.FlatGradSync(params)
    opt = torch.optim.Adam(params, lr=1e-3, capturable=True, fused=fused_adam)
    tr, _, _ = index_splits(series.size(0), 12)
    it = iter(IndexBatchLoader(series, tr, 12, 64, shuffle=True, seed=0))
    loss_fn = D.masked_mae_loss if fused_loss else D.masked_mae_loss_reference
    out = []
    for _ in range(steps):
        x, y = next(it)
        loss = loss_fn(head(model(x, ei, ew)[:, -1]).squeeze(-1), y[:, 0, :, 0])
        loss.backward()
        sync.all_reduce(); opt.step(); sync.zero()
        out.append(float(loss))
    return out


if __name__ == "__main__":
    for fl in (False, True):
        for fa in (False, True):
            print(f"fused_loss={int(fl)} fused_adam={int(fa)}", " ".join(f"{v:.6f}" for v in run(fl, fa)))
