#!/usr/bin/env python
"""BASELINE configs[2] (A3TGCN2 on the PEMS-BAY shape: 325 nodes, batch 64, 12 periods, 2 -> 32 channels): one training step
(forward + backward + Adam) through the fused kernel pair (stmp_tgcn_attn_fwd / _bwd) and through the op-for-op autograd path over
SpMM + cuBLAS, plus the inference launch.  One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pytorch_geometric_temporal_b200 import _lib  # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic  # noqa: E402
from pytorch_geometric_temporal_b200.nn.recurrent import A3TGCN2  # noqa: E402

dev = torch.device("cuda", 0)
ei, ew, _ = synthetic.pems_bay_like(0, 16)
ei, ew = torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev)
torch.manual_seed(0)
B, N, F, P = 64, 325, 2, 12
X = torch.randn(B, N, F, P, device=dev)
Y = torch.randn(B, N, P, device=dev)


def step_time(fused, iters=20):
    torch.manual_seed(1)
    m = A3TGCN2(F, 32, P, B).to(dev)
    m._base_tgcn.fused_training = fused
    head = torch.nn.Linear(32, P).to(dev)                      # examples/recurrent/a3tgcn2_example.py: relu -> Linear(32, periods)
    opt = torch.optim.Adam(list(m.parameters()) + list(head.parameters()), lr=1e-3)

    def body():
        opt.zero_grad(set_to_none=False)
        loss = torch.nn.functional.mse_loss(head(torch.relu(m(X, ei, ew))), Y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        body()
    n0 = _lib.launch_count()
    body()
    ours = _lib.launch_count() - n0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = body()
    e1.record()
    torch.cuda.synchronize()
    return {"ms_per_step": round(e0.elapsed_time(e1) / iters, 4), "our_launches_per_step": ours, "loss": float(loss)}


res = {"config": "A3TGCN2(2,32,periods=12) PEMS-BAY shape, batch 64, eager (no CUDA graph), fwd + bwd + Adam",
       "fused": step_time(True), "autograd_op_for_op": step_time(False)}
with torch.no_grad():
    m = A3TGCN2(F, 32, P, B).to(dev)
    for _ in range(3):
        m(X, ei, ew)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        m(X, ei, ew)
    e1.record()
    torch.cuda.synchronize()
    res["inference_us"] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
print(json.dumps(res))
