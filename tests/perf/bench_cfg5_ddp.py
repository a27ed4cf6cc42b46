#!/usr/bin/env python
"""BASELINE configs[4]: GConvLSTM(64,64,K=3) on the synthetic 10 000-node / 100 000-edge static graph, index-batched data-parallel
TRAINING with one flat NCCL gradient all-reduce per step (examples/indexBatching/DCRNN/pems_ddp.py:81-121 with the GConvLSTM cell).

  python tests/perf/bench_cfg5_ddp.py                                   (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/perf/bench_cfg5_ddp.py

Every rank keeps the whole normalised series (T_total x N x 64 floats = 655 MB) resident in HBM and draws its own shard of window
starts (DistributedSampler semantics); a step = 8 windows x 12 recurrent steps per GPU through the hand-written forward/backward
(`_LstmCellFn`), Linear head, masked-MAE loss, `FlatGradSync.all_reduce`, Adam.  Prints one JSON line (rank 0): snapshots/s over
all ranks, device-timed, max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pytorch_geometric_temporal_b200 import _lib, distributed as D           # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic                # noqa: E402
from pytorch_geometric_temporal_b200.nn.recurrent import GConvLSTM           # noqa: E402
from pytorch_geometric_temporal_b200.signal import IndexBatchLoader, index_splits  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=8)
    ap.add_argument("--nodes", type=int, default=10000)
    ap.add_argument("--edges", type=int, default=100000)
    ap.add_argument("--t-total", type=int, default=256)
    ap.add_argument("--autograd", action="store_true", help="op-for-op autograd path instead of the hand-written cell backward")
    args = ap.parse_args()
    rank, world, dev = D.init_process_group()
    N, Fd, Hd, K, HOR = args.nodes, 64, 64, 3, 12
    ei, ew = synthetic.large_graph(N, args.edges, 0)
    ei_d, ew_d = torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    series = torch.randn(args.t_total, N, Fd, device=dev, generator=g)
    torch.manual_seed(0)
    cell = GConvLSTM(Fd, Hd, K).to(dev)
    cell.fused_training = not args.autograd
    head = torch.nn.Linear(Hd, Fd).to(dev)
    params = list(cell.parameters()) + list(head.parameters())
    if world > 1:
        D.broadcast_parameters(cell); D.broadcast_parameters(head)
    sync = D.FlatGradSync(params)
    opt = torch.optim.Adam(params, lr=1e-3)
    tr, _, _ = index_splits(args.t_total, HOR)
    loader = IndexBatchLoader(series, tr, HOR, args.windows, shuffle=True, world_size=world, rank=rank, seed=0, drop_last=True)
    state = {"it": None, "epoch": 0}

    def next_batch():
        for _ in range(2):
            if state["it"] is None:
                loader.set_epoch(state["epoch"])
                state["it"] = iter(loader)
            try:
                return next(state["it"])
            except StopIteration:
                state["it"], state["epoch"] = None, state["epoch"] + 1
        raise RuntimeError("empty epoch")

    def step():
        x, y = next_batch()                                   # (B,12,N,64) each, gathered on the device from the resident series
        H = C = None
        for t in range(HOR):
            H, C = cell(x[:, t], ei_d, ew_d, H, C)
        loss = D.masked_mae_loss(head(H), y[:, 0])
        loss.backward()
        sync.all_reduce()
        opt.step()
        sync.zero()
        return loss.detach()

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if rank == 0:
        pc = {k: v for k, v in _lib.path_counters().items() if v and k in ("k_gemm_split", "k_lstm_gate_bwd", "k_spmm")}
        print(json.dumps({"config": f"cfg5 GConvLSTM(64,64,K=3) training, {N} nodes / {args.edges} edges, {args.windows} windows x 12 steps per GPU, "
                                    f"index-batched DDP, resident series {args.t_total}x{N}x64",
                          "n_gpus": world, "ms_per_step": ms, "snapshots_per_s": world * args.windows / (ms * 1e-3), "loss": float(loss),
                          "backward": "autograd (op-for-op)" if args.autograd else "hand-written (_LstmCellFn)",
                          "allreduce_bytes_per_step": sync.nbytes if world > 1 else 0, "stmp_launches_per_step": (_lib.launch_count() - l0) / args.steps,
                          "kernels": pc}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
