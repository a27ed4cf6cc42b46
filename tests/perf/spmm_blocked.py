#!/usr/bin/env python
"""SpMM at the cfg5 probe size: consecutive destination rows per lane group (stmp_set_option("spmm_rows_per_group")) on the random graph of
BASELINE configs[4] and on a banded (sensor-network-like) graph of the same size.  One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pytorch_geometric_temporal_b200 import _lib, ops  # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic  # noqa: E402
from pytorch_geometric_temporal_b200.plan import GraphPlan  # noqa: E402

dev = torch.device("cuda", 0)
B, N, F = 32, 10000, 128
x = torch.randn(B, N, F, device=dev)
y = torch.empty_like(x)
graphs = {"random": synthetic.large_graph(N, 100000, 0), "banded64": synthetic.banded_graph(N, 100000, 64, 0),
          "banded512": synthetic.banded_graph(N, 100000, 512, 0)}
res = {}
for name, (ei, ew) in graphs.items():
    plan = GraphPlan(_lib.FLAVOR_CHEB, torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev), N, normalization="sym")
    bytes_alg = B * (8 * N * F) + 8 * plan.nnz(0) + 4 * (N + 1)
    ref = None
    for blk, rpg in ((256, 1), (256, 4), (256, 8), (256, 12), (1024, 1), (1024, 2), (1024, 4), (1024, 8)):
        _lib.set_option("spmm_block", blk)
        _lib.set_option("spmm_rows_per_group", rpg)
        for _ in range(3):
            ops.spmm_raw(plan, 0, x, out=y)
        if ref is None:
            ref = y.clone()
        assert torch.equal(ref, y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            ops.spmm_raw(plan, 0, x, out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res.setdefault(name, {})[f"{blk}x{rpg}"] = {"ms": round(ms, 4), "gbs": round(bytes_alg / ms / 1e6, 1)}
_lib.set_option("spmm_block", 256)
_lib.set_option("spmm_rows_per_group", 8)
print(json.dumps(res))
