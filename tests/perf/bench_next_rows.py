#!/usr/bin/env python
"""Measurements for the SURVEY 8f rank-1 modules (GCLSTM, STConv, MSTGCN): ours (CUDA events after warm-up, CUDA-graph
replay where the launch sequence is fixed) next to the oracle port on the host cores.  One JSON object per line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_configs import DEV, cpu_time, gpu_time, graphed  # noqa: E402
from oracle import attention as OA, recurrent as R  # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic  # noqa: E402
from pytorch_geometric_temporal_b200.nn.attention import MSTGCN, STConv  # noqa: E402
from pytorch_geometric_temporal_b200.nn.recurrent import GCLSTM  # noqa: E402


def main():
    out = []
    with torch.no_grad():
        # GCLSTM at the cfg5 shape class (10k nodes / 100k edges, 64 hidden, K=3): same kernels as GConvLSTM
        ei, ew = synthetic.large_graph(10000, 100000, 0)
        ei, ew = torch.from_numpy(ei), torch.from_numpy(ew)
        torch.manual_seed(0)
        m = GCLSTM(64, 64, 3); sd = m.state_dict()
        X = torch.randn(8, 12, 10000, 64)
        mg, Xg, eig, ewg = m.to(DEV), X.to(DEV), ei.to(DEV), ew.to(DEV)
        def ours():
            H = C = None
            for t in range(12):
                H, C = mg(Xg[:, t], eig, ewg, H, C)
        def cpu():
            H = C = None
            for t in range(2):
                H, C = R.gc_lstm_cell(sd, X[0, t], ei, ew, H, C)
        run, how = graphed(ours)
        ms = gpu_time(run, iters=10)
        cms = cpu_time(cpu, budget=8) * 6 * 8
        out.append({"config": "GCLSTM(64,64,K=3), 10k nodes / 100k edges, 8 windows x 12 steps, forward", "launch": how, "ours_ms": ms,
                    "ours_snapshots_per_s": 8 / ms * 1e3, "cpu_oracle_ms_scaled": cms, "cpu_snapshots_per_s": 8 / cms * 1e3})
        # STConv on the PEMS-BAY shape: batch 32, 12 steps, 325 nodes, 2 -> 64 -> 64 channels, kernel 3, K=3
        ei, ew, _ = synthetic.pems_bay_like(0, 16)
        ei, ew = torch.from_numpy(ei), torch.from_numpy(ew)
        torch.manual_seed(0)
        m = STConv(325, 2, 64, 64, 3, 3).eval(); sd = m.state_dict()
        X = torch.randn(32, 12, 325, 2)
        mg, Xg, eig, ewg = m.to(DEV), X.to(DEV), ei.to(DEV), ew.to(DEV)
        run, how = graphed(lambda: mg(Xg, eig, ewg))
        ms = gpu_time(run)
        cms = cpu_time(lambda: OA.stconv(sd, X[:4], ei, ew, training=False), budget=8) * 8
        out.append({"config": "STConv(325 nodes, 2->64->64, kernel 3, K=3), batch 32 x 12 steps, forward (eval BatchNorm)", "launch": how,
                    "ours_ms": ms, "ours_windows_per_s": 32 / ms * 1e3, "cpu_oracle_ms_scaled_from_4_rows": cms,
                    "cpu_windows_per_s": 32 / cms * 1e3, "note": "reference loops ChebConv over 32*10 (b,t) slices; here one SpMM per hop"})
        # MSTGCN on the PeMS04 shape (cfg4's graph): 3 blocks, K=3, 64/64 filters, batch 32, 12 -> 12 steps
        eiu = torch.from_numpy(synthetic.pems04_like(0))
        torch.manual_seed(0)
        m = MSTGCN(3, 1, 3, 64, 64, 1, 12, 12); sd = m.state_dict()
        X = torch.randn(32, 307, 1, 12)
        mg, Xg, eg = m.to(DEV), X.to(DEV), eiu.to(DEV)
        mg(Xg, eg)                                            # lambda_max (host ARPACK) is computed once per static graph
        lam = mg._blocklist[0]._lambda_max(eg, 307)
        run, how = graphed(lambda: mg(Xg, eg))
        ms = gpu_time(run)
        cms = cpu_time(lambda: OA.mstgcn(sd, X[:4], eiu, 3, 1, lam), budget=8) * 8
        out.append({"config": "MSTGCN(3 blocks,K=3,64/64) PeMS04 shape (307 nodes), batch 32, forward", "launch": how, "ours_ms": ms,
                    "ours_windows_per_s": 32 / ms * 1e3, "cpu_oracle_ms_scaled_from_4_rows": cms, "cpu_windows_per_s": 32 / cms * 1e3,
                    "note": "oracle timed with lambda_max given; the reference also runs scipy ARPACK in every block forward"})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
