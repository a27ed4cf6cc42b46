#!/usr/bin/env python
"""Fused DCRNN forward at the reference's batch size (64 windows of 12 steps, METR-LA shape): one CTA per window vs a 2-CTA cluster per window.
One JSON line (microseconds per launch, inference and with the training stash)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pytorch_geometric_temporal_b200 import _lib  # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic  # noqa: E402
from pytorch_geometric_temporal_b200.nn.recurrent import BatchedDCRNN  # noqa: E402

dev = torch.device("cuda", 0)
ei, ew, _ = synthetic.metr_la_like(0, 16)
ei, ew = torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev)
torch.manual_seed(0)
m = BatchedDCRNN(2, 32, 2).to(dev)
res = {}
for B in (16, 64, 74):
    X = torch.randn(B, 12, 207, 2, device=dev)
    for split in (0, 1, 0, 1):
        _lib.set_option("dcrnn_fwd_split", split)
        with torch.no_grad():
            for _ in range(5):
                m(X, ei, ew)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                m(X, ei, ew)
            e1.record()
            torch.cuda.synchronize()
        res.setdefault(f"B{B}", {}).setdefault("cluster2" if split else "one_cta", []).append(round(e0.elapsed_time(e1) / 50 * 1e3, 1))
_lib.set_option("dcrnn_fwd_split", 0)
print(json.dumps(res))
