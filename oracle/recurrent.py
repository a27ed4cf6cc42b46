"""Functional CPU restatement of the reference's recurrent cells (oracle; test infrastructure).

Every function takes ``p``: a mapping with the reference module's ``state_dict()`` keys, so the same
tensors drive the oracle, the unmodified reference (in-container, via oracle/refload.py) and the CUDA
engine.  Reference = /root/reference/torch_geometric_temporal/nn/recurrent/*.py (cited per function).
Quirks of the reference are reproduced literally -- they are the spec.
"""
from typing import Mapping, Optional

import torch
from torch import Tensor

from . import pyg


def _sub(p: Mapping[str, Tensor], prefix: str) -> dict:
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


# ----------------------------------------------------------------------------------------------
# DConv / DCRNN                                                        nn/recurrent/dcrnn.py
# ----------------------------------------------------------------------------------------------
def dconv_operators(edge_index: Tensor, edge_weight: Optional[Tensor], batched: bool, num_nodes: int):
    """Returns (ei_out, norm_out, ei_in, norm_in): the two weighted edge lists DConv propagates over.

    unbatched (dcrnn.py:59-77): degrees from a dense adjacency (duplicates summed, sized by max index+1),
    both norms indexed by ``row``; reverse list = nonzero(adj^T) in row-major order, paired
    POSITIONALLY with norm_in.  Messages never use edge_weight (dcrnn.py:39-40).
    batched (dcrnn.py:277-290): degrees by scatter_add over X.size(0) nodes, reverse list =
    stack([col,row]) sorted by col*N+row.
    """
    row, col = edge_index[0], edge_index[1]
    if not batched:
        adj = pyg.to_dense_adj(edge_index, edge_attr=edge_weight)
        adj = adj.reshape(adj.size(1), adj.size(2))
        deg_out = torch.matmul(adj, torch.ones(adj.size(0), 1, device=adj.device)).flatten()
        deg_in = torch.matmul(torch.ones(1, adj.size(0), device=adj.device), adj).flatten()
        rev, _ = pyg.dense_to_sparse(adj.transpose(0, 1))
    else:
        deg_out = torch.zeros(num_nodes, device=row.device).scatter_add_(0, row, edge_weight)
        deg_in = torch.zeros(num_nodes, device=row.device).scatter_add_(0, col, edge_weight)
        rev = torch.stack([col, row], dim=0)
        rev = rev[:, (rev[0] * num_nodes + rev[1]).argsort()]
    norm_out = torch.reciprocal(deg_out)[row]
    norm_in = torch.reciprocal(deg_in)[row]
    return edge_index, norm_out, rev, norm_in


def dconv(X: Tensor, ops, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """dcrnn.py:79-111.  weight (2,K,C,out).  T_k = 2*prop(T_{k-1}) - X for every k>=2 (the reference
    never advances Tx_0 past X, :80,106)."""
    ei_o, n_o, ei_i, n_i = ops
    K = weight.size(1)
    H = torch.matmul(X, weight[0][0]) + torch.matmul(X, weight[1][0])
    if K > 1:
        To = pyg.propagate(ei_o, X, n_o)
        Ti = pyg.propagate(ei_i, X, n_i)
        H = H + torch.matmul(To, weight[0][1]) + torch.matmul(Ti, weight[1][1])
    for k in range(2, K):
        To = 2.0 * pyg.propagate(ei_o, To, n_o) - X
        Ti = 2.0 * pyg.propagate(ei_i, Ti, n_i) - X
        H = H + torch.matmul(To, weight[0][k]) + torch.matmul(Ti, weight[1][k])
    if bias is not None:
        H = H + bias
    return H


def _dcrnn_step(p, X, ops, H):
    """dcrnn.py:172-192."""
    g = lambda n, inp: dconv(inp, ops, p[f"conv_x_{n}.weight"], p.get(f"conv_x_{n}.bias"))
    Z = torch.sigmoid(g("z", torch.cat([X, H], dim=1)))
    R = torch.sigmoid(g("r", torch.cat([X, H], dim=1)))
    Ht = torch.tanh(g("h", torch.cat([X, H * R], dim=1)))
    return Z * H + (1 - Z) * Ht


def dcrnn_cell(p, X: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None, H: Optional[Tensor] = None):
    """DCRNN.forward (dcrnn.py:194-219); H None -> zeros (:167-170)."""
    out = p["conv_x_z.weight"].size(-1)
    if H is None:
        H = torch.zeros(X.shape[0], out, device=X.device)
    ops = dconv_operators(edge_index, edge_weight, batched=False, num_nodes=X.shape[0])
    return _dcrnn_step(p, X, ops, H)


def batched_dcrnn_operators(edge_index: Tensor, edge_weight: Tensor, B: int, N: int):
    """The cached part of BatchedDCRNN.forward (dcrnn.py:363-369,446-460): the block-diagonal edge list and, through
    BatchedDConv's `cached_idx`, its norms and sorted reverse list -- built once per (graph, batch size)."""
    ei = torch.cat([edge_index + i * N for i in range(B)], dim=1)
    ew = edge_weight.repeat(B)
    return dconv_operators(ei, ew, batched=True, num_nodes=B * N)


def batched_dcrnn(p, X: Tensor, edge_index: Tensor, edge_weight: Tensor, ops=None) -> Tensor:
    """BatchedDCRNN.forward (dcrnn.py:429-475): X (B,T,N,F) -> (B,T,N,out); graph replicated
    block-diagonally (:363-369), H0 = 0, Python loop over t.  `ops` = batched_dcrnn_operators(...) when the caller
    keeps them across calls, as the reference does (`cached_idx`, dcrnn.py:446-460).  Device-agnostic: tensors on
    `cuda` give the reference's op-for-op sequence on the GPU (bench.py's `reference_gpu` leg)."""
    B, T, N, F = X.shape
    out = p["conv_x_z.weight"].size(-1)
    if ops is None:
        ops = batched_dcrnn_operators(edge_index, edge_weight, B, N)
    H = torch.zeros(B * N, out, device=X.device)
    outs = []
    for t in range(T):
        H = _dcrnn_step(p, X[:, t].reshape(B * N, F), ops, H)
        outs.append(H.reshape(B, N, out))
    return torch.stack(outs, dim=1)


# ----------------------------------------------------------------------------------------------
# ChebConv cells                                     nn/recurrent/gconv_gru.py, gconv_lstm.py
# ----------------------------------------------------------------------------------------------
def cheb_conv(p, x, ei_norm, K: int):
    """PyG ChebConv given the already-normalised (edge_index', w_hat) pair; params 'lins.k.weight', 'bias'."""
    ei, w = ei_norm
    # .contiguous(): with a strided (permuted) input, ATen's CPU linear takes a different summation path for a
    # detached weight than for the reference's nn.Parameter (1 ulp apart); the Parameter path equals the contiguous one.
    lin = lambda k, t: torch.nn.functional.linear(t.contiguous(), p[f"lins.{k}.weight"])
    T0, T1 = x, x
    out = lin(0, T0)
    if K > 1:
        T1 = pyg.propagate(ei, x, w)
        out = out + lin(1, T1)
    for k in range(2, K):
        T2 = 2.0 * pyg.propagate(ei, T1, w) - T0
        out = out + lin(k, T2)
        T0, T1 = T1, T2
    if "bias" in p and p["bias"] is not None:
        out = out + p["bias"]
    return out


def _cheb_K(p, name):
    return len([k for k in p if k.startswith(name + ".lins.")])


def gconv_gru_cell(p, X, edge_index, edge_weight=None, H=None, lambda_max=None, normalization="sym"):
    """GConvGRU.forward (gconv_gru.py:141-170)."""
    K = _cheb_K(p, "conv_x_z")
    out = p["conv_x_z.lins.0.weight"].size(0)
    if H is None:
        H = torch.zeros(X.shape[0], out)
    en = pyg.cheb_norm(edge_index, X.size(-2), edge_weight, normalization, lambda_max, X.dtype)
    c = lambda name, t: cheb_conv(_sub(p, name + "."), t, en, K)
    Z = torch.sigmoid(c("conv_x_z", X) + c("conv_h_z", H))
    R = torch.sigmoid(c("conv_x_r", X) + c("conv_h_r", H))
    Ht = torch.tanh(c("conv_x_h", X) + c("conv_h_h", H * R))
    return Z * H + (1 - Z) * Ht


def gconv_lstm_cell(p, X, edge_index, edge_weight=None, H=None, C=None, lambda_max=None, normalization="sym"):
    """GConvLSTM.forward (gconv_lstm.py:204-238); O uses the NEW cell state (:235-236)."""
    K = _cheb_K(p, "conv_x_i")
    out = p["conv_x_i.lins.0.weight"].size(0)
    if H is None:
        H = torch.zeros(X.shape[0], out)
    if C is None:
        C = torch.zeros(X.shape[0], out)
    en = pyg.cheb_norm(edge_index, X.size(-2), edge_weight, normalization, lambda_max, X.dtype)
    c = lambda name, t: cheb_conv(_sub(p, name + "."), t, en, K)
    I = torch.sigmoid(c("conv_x_i", X) + c("conv_h_i", H) + p["w_c_i"] * C + p["b_i"])
    Fg = torch.sigmoid(c("conv_x_f", X) + c("conv_h_f", H) + p["w_c_f"] * C + p["b_f"])
    T = torch.tanh(c("conv_x_c", X) + c("conv_h_c", H) + p["b_c"])
    C = Fg * C + I * T
    O = torch.sigmoid(c("conv_x_o", X) + c("conv_h_o", H) + p["w_c_o"] * C + p["b_o"])
    return O * torch.tanh(C), C


def gc_lstm_cell(p, X, edge_index, edge_weight=None, H=None, C=None, lambda_max=None, normalization="sym"):
    """GCLSTM.forward (gc_lstm.py:152-205): dense `X @ W_g`, ChebConv only on H, no peepholes; O does not see C."""
    K = _cheb_K(p, "conv_i")
    out = p["W_i"].size(1)
    if H is None:
        H = torch.zeros(X.shape[0], out)
    if C is None:
        C = torch.zeros(X.shape[0], out)
    en = pyg.cheb_norm(edge_index, X.size(-2), edge_weight, normalization, lambda_max, X.dtype)
    c = lambda g: cheb_conv(_sub(p, f"conv_{g}."), H, en, K)
    I = torch.sigmoid(torch.matmul(X, p["W_i"]) + c("i") + p["b_i"])
    Fg = torch.sigmoid(torch.matmul(X, p["W_f"]) + c("f") + p["b_f"])
    T = torch.tanh(torch.matmul(X, p["W_c"]) + c("c") + p["b_c"])
    C = Fg * C + I * T
    O = torch.sigmoid(torch.matmul(X, p["W_o"]) + c("o") + p["b_o"])
    return O * torch.tanh(C), C


# ----------------------------------------------------------------------------------------------
# GCNConv cells                       nn/recurrent/temporalgcn.py, attentiontemporalgcn.py
# ----------------------------------------------------------------------------------------------
def gcn_conv(p, x, en):
    ei, w = en
    x = torch.nn.functional.linear(x, p["lin.weight"])
    out = pyg.propagate(ei, x, w)
    if p.get("bias") is not None:
        out = out + p["bias"]
    return out


def tgcn_cell(p, X, edge_index, edge_weight=None, H=None, improved=False, add_self_loops=True):
    """TGCN.forward (temporalgcn.py:104-130) and TGCN2.forward (:212-233): identical maths, the
    concatenation is over the last axis; X (N,F) or (B,N,F)."""
    out = p["conv_z.lin.weight"].size(0)
    if H is None:
        H = torch.zeros(*X.shape[:-1], out)
    en = pyg.gcn_norm(edge_index, edge_weight, X.size(-2), improved, add_self_loops, X.dtype)
    lin = lambda n, t: torch.nn.functional.linear(t, p[f"linear_{n}.weight"], p[f"linear_{n}.bias"])
    g = lambda n: gcn_conv(_sub(p, f"conv_{n}."), X, en)
    Z = torch.sigmoid(lin("z", torch.cat([g("z"), H], dim=-1)))
    R = torch.sigmoid(lin("r", torch.cat([g("r"), H], dim=-1)))
    Ht = torch.tanh(lin("h", torch.cat([g("h"), H * R], dim=-1)))
    return Z * H + (1 - Z) * Ht


def a3tgcn(p, X, edge_index, edge_weight=None, H=None, improved=False, add_self_loops=True):
    """A3TGCN.forward (attentiontemporalgcn.py:51-79) / A3TGCN2.forward (:130-157): X (...,N,F,P);
    the SAME H enters every period."""
    probs = torch.softmax(p["_attention"], dim=0)
    base = _sub(p, "_base_tgcn.")
    acc = 0
    for t in range(X.shape[-1]):
        acc = acc + probs[t] * tgcn_cell(base, X[..., t], edge_index, edge_weight, H, improved, add_self_loops)
    return acc
