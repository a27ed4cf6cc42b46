"""Host-side logic of the ChebConv-family modules (weight folding, gate order, shared Chebyshev basis, timestep folding,
the MSTGCN reshape, autograd through the torch part) checked on the CPU against the committed reference goldens.  The
two things that need a GPU -- the cached plan and `stmp_spmm` -- are replaced by a dense Laplacian built with the oracle's
`cheb_norm`; the modules run their autograd (training) code path, which is plain torch around those two calls."""
import os

import pytest
import torch

from oracle import pyg
from pytorch_geometric_temporal_b200 import ops
from pytorch_geometric_temporal_b200.nn.attention import MSTGCN, STConv
from pytorch_geometric_temporal_b200.nn.recurrent import GCLSTM, GConvGRU, GConvLSTM
import pytorch_geometric_temporal_b200.nn.recurrent._cheb as cheb_mod
import pytorch_geometric_temporal_b200.nn.recurrent.gc_lstm as gc_mod
import pytorch_geometric_temporal_b200.nn.recurrent.gconv_gru as gru_mod
import pytorch_geometric_temporal_b200.nn.recurrent.gconv_lstm as lstm_mod


class _DensePlan(object):
    def __init__(self, L):
        self.L = L


@pytest.fixture()
def dense_graph_ops(monkeypatch):
    def plan(self, edge_index, edge_weight, num_nodes, normalization, lambda_max, batch=None):
        lam = self._lambda_value(lambda_max)
        e, w = pyg.cheb_norm(edge_index, num_nodes, edge_weight, normalization, lam)
        L = torch.zeros(num_nodes, num_nodes)
        L.index_put_((e[1], e[0]), w, accumulate=True)              # out[dst] += w * x[src]
        return _DensePlan(L)

    def spmm(plan, op, x, alpha=1.0, z=None, beta=0.0, att=None):
        y = alpha * torch.matmul(plan.L, x)
        return y if z is None else y + beta * z

    monkeypatch.setattr(cheb_mod.ChebPlanMixin, "_cheb_plan", plan)
    monkeypatch.setattr(ops, "spmm", spmm)
    for m in (gc_mod, gru_mod, lstm_mod):
        monkeypatch.setattr(m, "_require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(GConvGRU, "_fused_ok", lambda self, plan, X, H: False)


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _loss(outs):
    return sum((o * torch.linspace(-1, 1, o.numel()).view_as(o)).sum() for o in outs)


def _close(a, b, rtol=1e-4, atol=1e-5):
    assert a.shape == b.shape and torch.allclose(a, b, rtol=rtol, atol=atol), float((a - b).abs().max())


def test_gconv_gru_and_lstm_host_logic(golden_dir, dense_graph_ops):
    g = _load(golden_dir, "gconv_gru_small")
    for c in g["cases"].values():
        m = GConvGRU(4, 16, c["K"], normalization=c["normalization"])
        m.load_state_dict(c["state"])
        _close(m(c["X"], g["edge_index"], g["edge_weight"], c["H"], c["lambda_max"]), c["out"])
    g = _load(golden_dir, "gconv_lstm_small")
    for c in g["cases"].values():
        m = GConvLSTM(4, 16, c["K"])
        m.load_state_dict(c["state"])
        h, cc = m(c["X"], g["edge_index"], g["edge_weight"], c["H"], c["C"])
        _close(h, c["outH"]); _close(cc, c["outC"])
        h, cc = m(c["X"], g["edge_index"])
        _close(h, c["outH0"]); _close(cc, c["outC0"])


def test_gc_lstm_host_logic_with_gradients(golden_dir, dense_graph_ops):
    g = _load(golden_dir, "gc_lstm_small")
    for name, c in g["cases"].items():
        if "grads" not in c:
            continue
        m = GCLSTM(4, 16, c["K"], normalization=c["normalization"])
        m.load_state_dict(c["state"])
        x, h, cc = (t.clone().requires_grad_(True) for t in (c["X"], c["H"], c["C"]))
        ho, co = m(x, g["edge_index"], g["edge_weight"], h, cc, c["lambda_max"])
        _close(ho, c["outH"]); _close(co, c["outC"])
        _loss([ho, co]).backward()
        for k, p in m.named_parameters():
            _close(p.grad, c["grads"][k], 1e-3, 1e-5)
        _close(x.grad, c["gX"], 1e-3, 1e-5); _close(h.grad, c["gH"], 1e-3, 1e-5); _close(cc.grad, c["gC"], 1e-3, 1e-5)


def test_stconv_host_logic_with_gradients(golden_dir, dense_graph_ops):
    g = _load(golden_dir, "stconv_small")
    for name, c in g["cases"].items():
        m = STConv(K=c["K"], normalization=c["normalization"], **g["ctor"])
        m.load_state_dict(c["state"])
        m.eval()
        with torch.no_grad():
            _close(m(c["X"], g["edge_index"], g["edge_weight"]), c["out_eval"])
            _close(m(c["X"], g["edge_index"]), c["out_eval_noew"])
        m.train()
        X = c["X"].clone().requires_grad_(True)
        out = m(X, g["edge_index"], g["edge_weight"])
        _close(out, c["out_train"])
        _loss([out]).backward()
        for k, p in m.named_parameters():                         # BatchNorm backward: bar relative to the tensor's max
            _close(p.grad, c["grads"][k], 0.0, 2e-4 * float(c["grads"][k].abs().max()) + 1e-6)
        _close(X.grad, c["gX"], 0.0, 2e-4 * float(c["gX"].abs().max()) + 1e-6)


def test_mstgcn_host_logic_with_gradients(golden_dir, dense_graph_ops):
    g = _load(golden_dir, "mstgcn_small")
    for name, c in g["cases"].items():
        m = MSTGCN(time_strides=c["time_strides"], **g["ctor"])
        m.load_state_dict(c["state"])
        X = c["X"].clone().requires_grad_(True)
        out = m(X, g["edge_index"])
        _close(out, c["out"], 2e-4, 2e-5)
        _loss([out]).backward()
        for k, p in m.named_parameters():
            _close(p.grad, c["grads"][k], 2e-3, 2e-5)
        _close(X.grad, c["gX"], 2e-3, 2e-5)
        with torch.no_grad():
            _close(m(c["X"], [g["edge_index"]] * 6), c["out_list"], 2e-4, 2e-5)


# ---- DConv / DCRNN (tiled path) and the GCNConv family ------------------------------------------------------------------
import pytorch_geometric_temporal_b200.nn.recurrent.attentiontemporalgcn as a3_mod  # noqa: E402
import pytorch_geometric_temporal_b200.nn.recurrent.dcrnn as dcrnn_mod  # noqa: E402
import pytorch_geometric_temporal_b200.nn.recurrent.temporalgcn as tgcn_mod  # noqa: E402
from oracle import recurrent as R  # noqa: E402
from pytorch_geometric_temporal_b200.nn.recurrent import A3TGCN, A3TGCN2, DCRNN, BatchedDCRNN, TGCN, TGCN2  # noqa: E402


class _DensePair(object):
    def __init__(self, mats):
        self.mats = mats


def _dense(edge_index, w, n):
    M = torch.zeros(n, n)
    M.index_put_((edge_index[1], edge_index[0]), w, accumulate=True)
    return M


@pytest.fixture()
def dense_dconv_gcn_ops(monkeypatch):
    def dconv_plan(self, edge_index, edge_weight, num_nodes):
        ew = edge_weight if edge_weight is not None else torch.ones(edge_index.size(1))
        eo, no, ei, ni = R.dconv_operators(edge_index, ew, self._batched_semantics, num_nodes)
        return _DensePair([_dense(eo, no, num_nodes), _dense(ei, ni, num_nodes)])

    def gcn_plan(self, edge_index, edge_weight, num_nodes):
        e, w = pyg.gcn_norm(edge_index, edge_weight, num_nodes, self.improved, self.add_self_loops, torch.float32)
        return _DensePair([_dense(e, w, num_nodes)])

    def spmm(plan, op, x, alpha=1.0, z=None, beta=0.0, att=None):
        y = alpha * torch.matmul(plan.mats[op], x)
        return y if z is None else y + beta * z

    monkeypatch.setattr(dcrnn_mod.DConv, "_plan", dconv_plan)
    monkeypatch.setattr(dcrnn_mod.DCRNN, "_plan", dconv_plan)
    monkeypatch.setattr(tgcn_mod.TGCN, "_plan", gcn_plan)
    monkeypatch.setattr(ops, "spmm", spmm)
    monkeypatch.setattr(ops, "dcrnn_seq_supported", lambda *a, **k: False)          # no fused kernels on the CPU
    monkeypatch.setattr(tgcn_mod.TGCN, "_fused_ok", lambda self, plan, X, H: False)
    for m in (dcrnn_mod, tgcn_mod, a3_mod):
        monkeypatch.setattr(m, "_require_cuda", lambda *a, **k: None)


def test_dcrnn_tiled_host_logic_with_gradients(golden_dir, dense_dconv_gcn_ops):
    for K in (1, 3, 4):
        g = _load(golden_dir, f"dcrnn_small_K{K}")
        m = DCRNN(3, 16, K)
        m.load_state_dict(g["state"])
        x, h = g["X"].clone().requires_grad_(True), g["H"].clone().requires_grad_(True)
        out = m(x, g["edge_index"], g["edge_weight"], h)
        _close(out, g["out"])
        _loss([out]).backward()
        for k, p in m.named_parameters():
            _close(p.grad, g["grads"][k], 1e-3, 1e-5)
        _close(x.grad, g["gX"], 1e-3, 1e-5); _close(h.grad, g["gH"], 1e-3, 1e-5)
    g = _load(golden_dir, "dcrnn_small_batched_K3")
    m = BatchedDCRNN(3, 16, 3)
    m.load_state_dict(g["state"])
    _close(m(g["X"], g["edge_index"], g["edge_weight"]), g["out"])
    g = _load(golden_dir, "dcrnn_cfg2_cell")
    m = DCRNN(2, 32, 2)
    m.load_state_dict(g["state"])
    _close(m(g["X"], g["edge_index"], g["edge_weight"], g["H"]), g["out"])
    _close(m(g["X"], g["edge_index"]), g["out_noew_noh"])                      # edge_weight None, H None


def test_tgcn_family_host_logic(golden_dir, dense_dconv_gcn_ops):
    g = _load(golden_dir, "tgcn_small")
    for c in g["cases"].values():
        m = TGCN(4, 16, improved=c["improved"], add_self_loops=c["add_self_loops"])
        m.load_state_dict(c["state"])
        _close(m(c["X"], g["edge_index"], g["edge_weight"], c["H"]), c["out"])
        m2 = TGCN2(4, 16, 3, improved=c["improved"], add_self_loops=c["add_self_loops"])
        m2.load_state_dict(c["state2"])
        _close(m2(c["X2"], g["edge_index"], g["edge_weight"], c["H2"]), c["out2"])
    g = _load(golden_dir, "a3tgcn_small")
    m = A3TGCN2(2, 16, 6, 3)
    m.load_state_dict(g["state"])
    _close(m(g["X"], g["edge_index"], g["edge_weight"]), g["out"])
    _close(m(g["X"], g["edge_index"], g["edge_weight"], torch.ones(3, 40, 16) * 0.3), g["outH"])
    m1 = A3TGCN(2, 16, 6)
    m1.load_state_dict(g["state1"])
    _close(m1(g["X1"], g["edge_index"], g["edge_weight"]), g["out1"])


def test_tgcn_attention_folding_host_logic_vs_reference_golden(golden_dir, dense_dconv_gcn_ops, monkeypatch):
    """Host side of the fused temporal-attention + GCN kernel (stmp_tgcn_attn_fwd): the folded weights (A, Bm, c) of
    `TGCN._packed3`, the period softmax and the call shapes, with the kernel replaced by a dense restatement of its
    arithmetic -- against the UNMODIFIED reference modules at the PEMS-BAY shape (325 nodes)."""
    def fake_kernel(plan, x, A, Bm, c, probs=None, h=None, h_shared=False):
        B, N, Fi, P = x.shape
        L = plan.mats[0]
        H = torch.zeros(B, N, 32) if h is None else (h.expand(B, N, 32) if h_shared else h.reshape(B, N, 32))
        out = torch.zeros(B, N, 32)
        for t in range(P):
            ax = torch.matmul(L, x[..., t])                                     # (B,N,Fi): A^ X_t
            pre = lambda g, Hp: ax @ A[:, 32 * g:32 * g + 32] + Hp @ Bm[:, 32 * g:32 * g + 32] + c[32 * g:32 * g + 32]
            Z, Rg = torch.sigmoid(pre(0, H)), torch.sigmoid(pre(1, H))
            Hn = Z * H + (1 - Z) * torch.tanh(pre(2, H * Rg))
            out = out + (Hn if probs is None else probs[t] * Hn)
        return out
    monkeypatch.setattr(ops, "tgcn_attn_fwd", fake_kernel)
    g = _load(golden_dir, "a3tgcn2_cfg3")
    ei, ew, X, H = g["edge_index"], g["edge_weight"], g["X"], g["H"]
    with torch.no_grad():
        m = A3TGCN2(2, 32, 12, 64)
        m.load_state_dict(g["state"])
        _close(m(X, ei, ew), g["out"]); _close(m(X, ei, ew, H), g["outH"])
        m1 = A3TGCN(2, 32, 12)
        m1.load_state_dict(g["state1"])
        _close(m1(X[0], ei, ew), g["out1"]); _close(m1(X[0], ei, ew, H[0]), g["out1H"])
        c2 = TGCN2(2, 32, 8)
        c2.load_state_dict(g["state_cell"])
        _close(c2(X[..., 0], ei, ew), g["cell"]); _close(c2(X[..., 0], ei, ew, H), g["cellH"])


def test_tgcn_attention_training_host_logic_vs_reference_golden_gradients(golden_dir, dense_dconv_gcn_ops, monkeypatch):
    """Host side of the fused TRAINING path (stmp_tgcn_attn_fwd + stmp_tgcn_attn_bwd): the differentiable folding `TGCN._fold3`, the
    softmax of the attention and the routing (no state, no input gradient -> `ops.tgcn_attn_train`), with the kernel pair replaced by a
    dense differentiable restatement -- output and EVERY parameter gradient against the unmodified reference at the PEMS-BAY shape."""
    calls = []

    def fake_train(plan, x, A, Bm, c, probs=None):
        calls.append(tuple(x.shape))
        L = plan.mats[0]
        out = 0
        for t in range(x.shape[-1]):
            ax = torch.matmul(L, x[..., t])
            Z = torch.sigmoid(ax @ A[:, 0:32] + c[0:32])
            Ht = torch.tanh(ax @ A[:, 64:96] + c[64:96])                       # H = 0: the r gate and Bm drop out
            out = out + (1 - Z) * Ht * (1.0 if probs is None else probs[t])
        return out
    monkeypatch.setattr(ops, "tgcn_attn_train", fake_train)
    g = _load(golden_dir, "a3tgcn2_cfg3_grads")
    ei, ew, X = g["edge_index"], g["edge_weight"], g["X"]
    m = A3TGCN2(2, 32, 12, 8)
    m.load_state_dict(g["state"])
    out = m(X, ei, ew)
    w = torch.linspace(-1, 1, out.numel()).view_as(out)
    (out * w).sum().backward()
    _close(out, g["out"])
    for k, p in m.named_parameters():
        ref = g["grads"][k]
        assert p.grad is not None, k
        assert torch.allclose(p.grad, ref, rtol=1e-3, atol=1e-3 * float(ref.abs().max()) + 1e-6), k
    c2 = TGCN2(2, 32, 8)
    c2.load_state_dict(g["state_cell"])
    cell = c2(X[..., 3], ei, ew)
    (cell * w).sum().backward()
    _close(cell, g["cell"])
    for k, p in c2.named_parameters():
        ref = g["grads_cell"][k]
        assert torch.allclose(p.grad, ref, rtol=1e-3, atol=1e-3 * float(ref.abs().max()) + 1e-6), k
    assert calls == [(8, 325, 2, 12), (8, 325, 2, 1)]
    # a call with an incoming state stays on the op-for-op path
    m(X[:2], ei, ew, torch.zeros(2, 325, 32)).sum().backward()
    assert len(calls) == 2


# ---- ASTGCN: attention-weighted first hop, timesteps folded into the feature axis, fp32 time convolutions ---------------
import pytorch_geometric_temporal_b200.nn.attention.astgcn as astgcn_mod  # noqa: E402
from oracle import attention as OA  # noqa: E402
from pytorch_geometric_temporal_b200.nn.attention import ASTGCN  # noqa: E402


@pytest.fixture()
def dense_att_ops(monkeypatch):
    def plan(self, edge_index, edge_weight, num_nodes, lambda_max, batch=None):
        lam = torch.tensor(2.0 if lambda_max is None else float(lambda_max))
        ei, w = OA.cheb_att_norm(edge_index, num_nodes, edge_weight, self._normalization, lam)
        W = torch.zeros(num_nodes, num_nodes)
        W.index_put_((ei[0], ei[1]), w, accumulate=True)          # propagated on the TRANSPOSED index: out[row] += w x[col]
        return _DensePlan(W)

    def spmm(plan, op, x, alpha=1.0, z=None, beta=0.0, att=None):
        A = plan.L if att is None else plan.L.unsqueeze(0) * att  # first hop: norm * S[b,row,col]
        y = alpha * torch.matmul(A, x)
        return y if z is None else y + beta * z

    monkeypatch.setattr(astgcn_mod.ChebConvAttention, "_plan", plan)
    monkeypatch.setattr(ops, "spmm", spmm)
    monkeypatch.setattr(astgcn_mod, "_require_cuda", lambda *a, **k: None)


def test_astgcn_host_logic(golden_dir, dense_att_ops):
    g = _load(golden_dir, "astgcn_small")
    for name, c in g["cases"].items():
        m = ASTGCN(**g["ctor"], normalization=c["normalization"])
        m.load_state_dict(c["state"])
        X = c["X"].clone().requires_grad_(True)
        out = m(X, g["edge_index"])
        _close(out, c["out"], 2e-4, 2e-5)
        out.sum().backward()                                      # the whole block is differentiable through the stand-ins
        assert torch.isfinite(X.grad).all() and all(torch.isfinite(p.grad).all() for p in m.parameters())
        with torch.no_grad():
            _close(m(c["X"], [g["edge_index"]] * 6), c["out"], 2e-4, 2e-5)      # per-timestep list of the same graph
