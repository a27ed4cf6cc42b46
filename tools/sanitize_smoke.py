"""One small invocation of every hand-synchronised kernel (mbarriers, TMEM alloc, proxy fences, last-arriver MMA issue, TMA bulk copies)
for compute-sanitizer:   tools/sanitize.sh   runs this under --tool memcheck and --tool racecheck and keeps the logs in profiles/."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_temporal_b200 import _lib, ops                                        # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic                               # noqa: E402
from pytorch_geometric_temporal_b200.nn.attention import ASTGCN                              # noqa: E402
from pytorch_geometric_temporal_b200.nn.recurrent import A3TGCN2, BatchedDCRNN, GConvGRU, GConvLSTM   # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
ei, ew, series = synthetic.metr_la_like(0, 40)
ei_t, ew_t = torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev)
X = torch.from_numpy(series[:36]).reshape(3, 12, 207, 2).to(dev)
m = BatchedDCRNN(2, 32, 2).to(dev)
with torch.no_grad():
    m(X, ei_t, ew_t)                                             # k_dcrnn_seq_tc (tcgen05, TMA images, MMA groups)
m(X[:2], ei_t, ew_t).square().mean().backward()                  # + stash; CTA-pair (cluster) forward and backward, k_dcrnn_bwd_basis, k_dcrnn_wgrad_tc
for opt in ("dcrnn_fwd_split", "dcrnn_bwd_split", "dcrnn_wgrad_tc"):
    _lib.set_option(opt, 0)
with torch.no_grad():
    m(X, ei_t, ew_t)                                             # one CTA per window
m(X[:2], ei_t, ew_t).square().mean().backward()                  # one-CTA backward, FFMA weight-gradient kernel
for opt in ("dcrnn_fwd_split", "dcrnn_bwd_split", "dcrnn_wgrad_tc"):
    _lib.set_option(opt, 1)
from pytorch_geometric_temporal_b200 import distributed as D   # noqa: E402
sync = D.FlatGradSync(m.parameters())
opt_ = D.FlatAdam(sync, lr=1e-3)
m(X[:2], ei_t, ew_t).square().mean().backward()
opt_.step()                                                      # k_adam_flat
with torch.no_grad():
    BatchedDCRNN(2, 16, 3).to(dev)(X[:2], ei_t, ew_t)            # k_dcrnn_seq (FFMA2, TMA window buffers)
    g = GConvGRU(2, 32, 2).to(dev)
    g(X[0, 0], ei_t, ew_t)                                       # generic graph-GRU entry (one operator)
    e3, w3, _ = synthetic.pems_bay_like(0, 16)
    a3 = A3TGCN2(2, 32, 12, 4).to(dev)
    a3(torch.randn(4, 325, 2, 12, device=dev), torch.from_numpy(e3).to(dev), torch.from_numpy(w3).to(dev),
       torch.randn(4, 325, 32, device=dev))                      # k_tgcn_attn (TMA-staged X)
with torch.enable_grad():
    a3(torch.randn(4, 325, 2, 12, device=dev), torch.from_numpy(e3).to(dev), torch.from_numpy(w3).to(dev)).square().mean().backward()   # k_tgcn_attn_bwd
with torch.no_grad():
    e4 = torch.from_numpy(synthetic.pems04_like(0)).to(dev)
    ASTGCN(2, 1, 3, 64, 64, 1, 12, 12, 307, normalization="sym").to(dev)(torch.randn(2, 307, 1, 12, device=dev), e4)   # k_gemm_blocks x7
    eg, wg = synthetic.large_graph(2000, 20000, 0)
    eg, wg = torch.from_numpy(eg).to(dev), torch.from_numpy(wg).to(dev)
    lstm = GConvLSTM(64, 64, 3).to(dev)
    lstm(torch.randn(2000, 64, device=dev), eg, wg)              # k_spmm + k_gemm_split<LSTM epilogue>
x = torch.randn(2, 2000, 64, device=dev, requires_grad=True)
h, c = lstm(x, eg, wg)
(h.sum() + c.sum()).backward()                                   # _LstmCellFn backward: k_gemm_split, k_lstm_gate_bwd, transposed SpMM
torch.cuda.synchronize()
print("sanitize_smoke ok:", {k: v for k, v in _lib.path_counters().items() if k.startswith("k_") and v})
