"""Restatement of the torch_geometric primitives used on the hot path (oracle; test infrastructure).

PyG is an un-pinned third-party dependency of the reference (/root/reference/setup.py:7) and is not
vendored, so these follow PyG's published semantics (PyG 2.5/2.6, the versions the reference CI
resolves, .github/workflows/main.yml:31-32) and are anchored on the reference's call sites, cited per
function.  Float semantics here are "parity unpinned" (see oracle/__init__.py).
"""
import inspect
import math
from typing import Optional

import torch
from torch import Tensor


# --------------------------------------------------------------------------------------
# scatter / propagate  (call sites: nn/recurrent/dcrnn.py:86-87,95-99,300-313; nn/attention/astgcn.py:169-175)
# --------------------------------------------------------------------------------------
def scatter_add(src: Tensor, index: Tensor, dim: int, dim_size: int) -> Tensor:
    """PyG ``scatter(..., reduce='sum')``: zeros(size).scatter_add_(dim, broadcast(index), src)."""
    dim = dim if dim >= 0 else src.dim() + dim
    size = list(src.shape)
    size[dim] = dim_size
    shape = [1] * src.dim()
    shape[dim] = -1
    idx = index.view(shape).expand_as(src)
    return src.new_zeros(size).scatter_add_(dim, idx, src)


def propagate(edge_index: Tensor, x: Tensor, norm: Optional[Tensor]) -> Tensor:
    """MessagePassing.propagate, aggr='add', flow='source_to_target', node_dim=-2, with the
    message ``norm.view(-1,1) * x_j`` (dcrnn.py:39-40) or ``norm.view(d1,d2,1) * x_j`` for a
    per-batch norm (astgcn.py:185-190)."""
    n = x.size(-2)
    x_j = x.index_select(-2, edge_index[0])
    if norm is None:
        msg = x_j
    elif norm.dim() == 1:
        msg = norm.view(-1, 1) * x_j
    else:
        d1, d2 = norm.shape
        msg = norm.view(d1, d2, 1) * x_j
    return scatter_add(msg, edge_index[1], -2, n)


class MessagePassing(torch.nn.Module):
    """Minimal MessagePassing: only what the reference's subclasses use (aggr='add',
    flow='source_to_target', propagate(edge_index, size=None, **kwargs) -> message(...))."""

    def __init__(self, aggr: str = "add", flow: str = "source_to_target", node_dim: int = -2, **kwargs):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target"
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index: Tensor, size=None, **kwargs) -> Tensor:
        params = list(inspect.signature(self.message).parameters)
        ref = None
        args = {}
        for p in params:
            if p.endswith("_j") or p.endswith("_i"):
                data = kwargs[p[:-2]]
                ref = data
                sel = edge_index[0] if p.endswith("_j") else edge_index[1]
                args[p] = data.index_select(self.node_dim, sel)
            else:
                args[p] = kwargs[p]
        msg = self.message(**args)
        return scatter_add(msg, edge_index[1], self.node_dim, ref.size(self.node_dim))

    def message(self, x_j):  # pragma: no cover - overridden
        return x_j


# --------------------------------------------------------------------------------------
# dense <-> sparse  (dcrnn.py:59-60,76-77; dataset/metr_la.py:92)
# --------------------------------------------------------------------------------------
def to_dense_adj(edge_index: Tensor, batch=None, edge_attr: Optional[Tensor] = None, max_num_nodes=None) -> Tensor:
    """(1, N', N') with N' = edge_index.max()+1; duplicate edges SUM; edge_attr None -> ones."""
    n = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    if max_num_nodes is not None:
        n = max_num_nodes
    if edge_attr is None:
        edge_attr = torch.ones(edge_index.size(1), device=edge_index.device)
    flat = edge_index[0] * n + edge_index[1]
    adj = scatter_add(edge_attr, flat, 0, n * n)
    return adj.view(1, n, n)


def dense_to_sparse(adj: Tensor):
    """2-D adj: index = adj.nonzero().t() (row-major order), values = adj[index]."""
    assert adj.dim() == 2
    index = adj.nonzero().t().contiguous()
    return index, adj[index[0], index[1]]


# --------------------------------------------------------------------------------------
# self loops / laplacian / gcn_norm
# --------------------------------------------------------------------------------------
def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None, fill_value=1.0, num_nodes: Optional[int] = None):
    """Appends N loops (node order) after the existing edges (astgcn.py:104-106)."""
    n = num_nodes if num_nodes is not None else int(edge_index.max()) + 1
    loop = torch.arange(n, device=edge_index.device, dtype=edge_index.dtype)
    loop = loop.unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        loop_attr = edge_attr.new_full((n,), fill_value)
        edge_attr = torch.cat([edge_attr, loop_attr], dim=0)
    return torch.cat([edge_index, loop], dim=1), edge_attr


def add_remaining_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None, fill_value=1.0, num_nodes: Optional[int] = None):
    """Non-loop edges kept in order, then N loops in node order whose weight is the existing loop's
    weight where one was present, else ``fill_value`` (gcn_norm)."""
    n = num_nodes if num_nodes is not None else int(edge_index.max()) + 1
    mask = edge_index[0] != edge_index[1]
    loop = torch.arange(n, device=edge_index.device, dtype=edge_index.dtype).unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        loop_attr = edge_attr.new_full((n,), fill_value)
        inv = ~mask
        loop_attr[edge_index[0][inv]] = edge_attr[inv]
        edge_attr = torch.cat([edge_attr[mask], loop_attr], dim=0)
    return torch.cat([edge_index[:, mask], loop], dim=1), edge_attr


def get_laplacian(edge_index: Tensor, edge_weight: Optional[Tensor] = None, normalization: Optional[str] = None,
                  dtype=None, num_nodes: Optional[int] = None):
    """L = D - A (None), I - D^-1/2 A D^-1/2 ('sym'), I - D^-1 A ('rw'); loops appended AFTER the
    non-loop edges, node order."""
    assert normalization in (None, "sym", "rw")
    edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    n = num_nodes if num_nodes is not None else int(edge_index.max()) + 1
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, row, 0, n)
    if normalization is None:
        edge_index, _ = add_self_loops(edge_index, num_nodes=n)
        edge_weight = torch.cat([-edge_weight, deg], dim=0)
    elif normalization == "sym":
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float("inf"), 0)
        edge_weight = dis[row] * edge_weight * dis[col]
        edge_index, edge_weight = add_self_loops(edge_index, -edge_weight, fill_value=1.0, num_nodes=n)
    else:
        dinv = 1.0 / deg
        dinv.masked_fill_(dinv == float("inf"), 0)
        edge_weight = dinv[row] * edge_weight
        edge_index, edge_weight = add_self_loops(edge_index, -edge_weight, fill_value=1.0, num_nodes=n)
    return edge_index, edge_weight


def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int, improved: bool = False,
             add_self_loops_: bool = True, dtype=torch.float32):
    fill = 2.0 if improved else 1.0
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    if add_self_loops_:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, col, 0, num_nodes)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


# --------------------------------------------------------------------------------------
# inits (gconv_lstm.py:6,149-156)
# --------------------------------------------------------------------------------------
def glorot(t: Tensor):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


def zeros(t: Tensor):
    with torch.no_grad():
        t.fill_(0)


class _Linear(torch.nn.Module):
    """PyG ``Linear(in, out, bias=False, weight_initializer='glorot')``; state_dict key 'weight' (out,in)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        glorot(self.weight)

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight)


# --------------------------------------------------------------------------------------
# ChebConv (ctor sites gconv_gru.py:57-107, gconv_lstm.py:62-138)
# --------------------------------------------------------------------------------------
def cheb_norm(edge_index, num_nodes, edge_weight, normalization, lambda_max=None, dtype=torch.float32, batch=None):
    """Current-PyG ChebConv.__norm__: get_laplacian -> lambda_max None => 2*w.max() -> 2w/lam ->
    inf->0 -> loop entries -= 1."""
    edge_index, edge_weight = get_laplacian(edge_index, edge_weight, normalization, dtype, num_nodes)
    if lambda_max is None:
        lambda_max = 2.0 * edge_weight.max()
    elif not isinstance(lambda_max, Tensor):
        lambda_max = torch.tensor(lambda_max, dtype=dtype, device=edge_index.device)
    if batch is not None and lambda_max.numel() > 1:
        lambda_max = lambda_max[batch[edge_index[0]]]
    edge_weight = (2.0 * edge_weight) / lambda_max
    edge_weight.masked_fill_(edge_weight == float("inf"), 0)
    loop_mask = edge_index[0] == edge_index[1]
    edge_weight[loop_mask] -= 1
    return edge_index, edge_weight


class ChebConv(MessagePassing):
    def __init__(self, in_channels: int, out_channels: int, K: int, normalization: Optional[str] = "sym",
                 bias: bool = True, **kwargs):
        super().__init__(aggr="add")
        assert K > 0
        assert normalization in (None, "sym", "rw")
        self.in_channels, self.out_channels, self.normalization = in_channels, out_channels, normalization
        self.lins = torch.nn.ModuleList([_Linear(in_channels, out_channels) for _ in range(K)])
        if bias:
            self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def message(self, x_j, norm):
        return norm.view(-1, 1) * x_j

    def forward(self, x, edge_index, edge_weight=None, batch=None, lambda_max=None):
        edge_index, norm = cheb_norm(edge_index, x.size(self.node_dim), edge_weight, self.normalization,
                                     lambda_max, dtype=x.dtype, batch=batch)
        Tx_0 = x
        Tx_1 = x
        out = self.lins[0](Tx_0)
        if len(self.lins) > 1:
            Tx_1 = self.propagate(edge_index, x=x, norm=norm, size=None)
            out = out + self.lins[1](Tx_1)
        for lin in self.lins[2:]:
            Tx_2 = self.propagate(edge_index, x=Tx_1, norm=norm, size=None)
            Tx_2 = 2.0 * Tx_2 - Tx_0
            out = out + lin(Tx_2)
            Tx_0, Tx_1 = Tx_1, Tx_2
        if self.bias is not None:
            out = out + self.bias
        return out


# --------------------------------------------------------------------------------------
# GCNConv (ctor sites temporalgcn.py:38-68,162-173)
# --------------------------------------------------------------------------------------
class GCNConv(MessagePassing):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__(aggr="add")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached, self.add_self_loops, self.normalize = improved, cached, add_self_loops, normalize
        self.lin = _Linear(in_channels, out_channels)
        if bias:
            self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def message(self, x_j, edge_weight):
        return x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j

    def forward(self, x, edge_index, edge_weight=None):
        if self.normalize:
            edge_index, edge_weight = gcn_norm(edge_index, edge_weight, x.size(self.node_dim), self.improved,
                                               self.add_self_loops, x.dtype)
        x = self.lin(x)
        out = self.propagate(edge_index, x=x, edge_weight=edge_weight, size=None)
        if self.bias is not None:
            out = out + self.bias
        return out


# --------------------------------------------------------------------------------------
# containers / transforms (signal/static_graph_temporal_signal.py:4,119; astgcn.py:9,12,434-438)
# --------------------------------------------------------------------------------------
class Data(object):
    """Attribute bag standing in for torch_geometric.data.Data."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        if "_num_nodes" in self.__dict__:
            return self.__dict__["_num_nodes"]
        if self.x is not None:
            return self.x.size(0)
        return int(self.edge_index.max()) + 1

    @num_nodes.setter
    def num_nodes(self, v):
        self.__dict__["_num_nodes"] = v


class LaplacianLambdaMax(object):
    """scipy ARPACK largest eigenvalue of the (normalised) Laplacian (astgcn.py:437-438)."""

    def __init__(self, normalization=None, is_undirected=False):
        self.normalization, self.is_undirected = normalization, is_undirected

    def __call__(self, data):
        import numpy as np
        from scipy.sparse import coo_matrix
        from scipy.sparse.linalg import eigs, eigsh
        n = data.num_nodes
        ew = data.edge_attr
        ei, w = get_laplacian(data.edge_index, ew, self.normalization, num_nodes=n)
        L = coo_matrix((w.detach().cpu().numpy().astype(np.float64), (ei[0].cpu().numpy(), ei[1].cpu().numpy())), shape=(n, n))
        fn = eigs
        if self.is_undirected and self.normalization != "rw":
            fn = eigsh
        lam = fn(L, k=1, which="LM", return_eigenvectors=False)
        data.lambda_max = float(lam.real[0] if hasattr(lam, "real") else lam[0])
        return data
