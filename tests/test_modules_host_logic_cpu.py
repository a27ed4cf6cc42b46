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
    def plan(self, edge_index, edge_weight, num_nodes, normalization, lambda_max):
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
