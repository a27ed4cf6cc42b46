#!/usr/bin/env python
"""Run ONE of the secondary configs a few times (for `ncu --metrics gpu__time_duration.sum` launch lists / `ncu --set full` captures):
   python tests/perf/run_config.py cfg3|cfg4|cfg5 [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pytorch_geometric_temporal_b200.dataset import synthetic                                   # noqa: E402
from pytorch_geometric_temporal_b200.nn.attention import ASTGCN                                  # noqa: E402
from pytorch_geometric_temporal_b200.nn.recurrent import A3TGCN2, GConvLSTM                      # noqa: E402

cfg = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda")
torch.manual_seed(0)
with torch.no_grad():
    if cfg == "cfg3":
        ei, ew, _ = synthetic.pems_bay_like(0, 16)
        m = A3TGCN2(2, 32, 12, 64).to(dev)
        X, ei, ew = torch.randn(64, 325, 2, 12, device=dev), torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev)
        fn = lambda: m(X, ei, ew)
    elif cfg == "cfg4":
        ei = torch.from_numpy(synthetic.pems04_like(0)).to(dev)
        m = ASTGCN(3, 1, 3, 64, 64, 1, 12, 12, 307, normalization="sym").to(dev)
        X = torch.randn(32, 307, 1, 12, device=dev)
        fn = lambda: m(X, ei)
    else:
        ei, ew = synthetic.large_graph(10000, 100000, 0)
        ei, ew = torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev)
        m = GConvLSTM(64, 64, 3).to(dev)
        X = torch.randn(8, 12, 10000, 64, device=dev)

        def fn():
            H = C = None
            for t in range(12):
                H, C = m(X[:, t], ei, ew, H, C)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
print(cfg, "eager ms per call:", e0.elapsed_time(e1) / iters)
