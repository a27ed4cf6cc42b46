"""Time the SpMM kernel variants on the cfg5 probe (N=10^4, nnz=110 000, F=128, batch 32) and check they are bit-identical.
  python tests/perf/spmm_variants.py            (GPU box)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pytorch_geometric_temporal_b200 import _lib, ops                      # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic             # noqa: E402
from pytorch_geometric_temporal_b200.plan import GraphPlan                # noqa: E402


def main():
    dev = torch.device("cuda")
    ei, ew = synthetic.large_graph(10000, 100000, 0)
    plan = GraphPlan(_lib.FLAVOR_CHEB, torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev), 10000, normalization="sym")
    nnz = plan.nnz(0)
    res = {}
    for (B, F) in ((32, 128), (32, 64), (96, 64), (8, 384)):
        x = torch.randn(B, 10000, F, device=dev)
        ref = None
        for v in (0, 1, 2):
            _lib.set_option("spmm_variant", v)
            y = torch.empty_like(x)
            for _ in range(3):
                ops.spmm_raw(plan, 0, x, out=y)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                ops.spmm_raw(plan, 0, x, out=y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            alg = B * 8 * 10000 * F + 8 * nnz + 4 * 10001
            same = True if ref is None else bool(torch.equal(ref, y))
            ref = y.clone() if ref is None else ref
            res[f"B{B}_F{F}_v{v}"] = {"ms": round(ms, 4), "GBs": round(alg / ms / 1e6, 1), "bit_identical_to_v0": same}
    _lib.set_option("spmm_variant", 0)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
