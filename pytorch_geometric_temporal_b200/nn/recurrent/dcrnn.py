"""DCRNN family behind the reference's module surface (drop-in for
torch_geometric_temporal/nn/recurrent/dcrnn.py: DConv :7-111, DCRNN :114-219, BatchedDConv :222-325,
BatchedDCRNN :328-475).  Same constructor signatures, forward signatures and state_dict keys
(`conv_x_{z,r,h}.weight (2,K,Cin+Cout,Cout)`, `.bias (Cout)`); the arithmetic runs in libstmp:

* inference (no grad): the whole recurrence in ONE fused kernel (`stmp_dcrnn_seq_fwd`);
* training / shapes the fused kernel cannot take: the tiled path = hand-written SpMM (`stmp_spmm`,
  differentiable through its transposed product) + cuBLAS contraction, with the diffusion shared
  between the z and r gates.
"""
import torch

from ... import ops
from ... import _lib
from ...plan import PlanCache, _require_cuda


def _basis(plan, U: torch.Tensor, K: int):
    """[U, P_o U, P_i U, 2 P_o T - U, ...] for U (N,C) or (B,N,C); T_k = 2 P T_{k-1} - U for every
    k >= 2 (the reference never advances Tx_0 past X, dcrnn.py:80,106)."""
    blocks = [U]
    To = Ti = None
    for k in range(1, K):
        if k == 1:
            To, Ti = ops.spmm(plan, 0, U), ops.spmm(plan, 1, U)
        else:
            To = ops.spmm(plan, 0, To, alpha=2.0, z=U, beta=-1.0)
            Ti = ops.spmm(plan, 1, Ti, alpha=2.0, z=U, beta=-1.0)
        blocks += [To, Ti]
    return blocks


def _stack_weight(weight: torch.Tensor) -> torch.Tensor:
    """(2,K,C,O) -> ((2K-1)*C, O) matching `_basis` block order; block 0 = W[0,0] + W[1,0]."""
    K = weight.size(1)
    parts = [weight[0, 0] + weight[1, 0]]
    for k in range(1, K):
        parts += [weight[0, k], weight[1, k]]
    return torch.cat(parts, dim=0)


class DConv(torch.nn.Module):
    r"""Diffusion convolution (reference: dcrnn.py:7-111).  Messages use only the degree norms, never
    edge_weight (:39-40); norm_in is indexed by `row` and paired positionally with the re-sorted
    reverse edge list (:74-77) -- reproduced inside the plan (csrc/plan.cu)."""

    _batched_semantics = False

    def __init__(self, in_channels, out_channels, K, bias=True):
        super().__init__()
        assert K > 0
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight = torch.nn.Parameter(torch.empty(2, K, in_channels, out_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self._plans = PlanCache()
        self._reset_parameters()

    def _reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def _plan(self, edge_index, edge_weight, num_nodes):
        flags = _lib.DCONV_ALLOW_DUPLICATES if self._batched_semantics else 0
        return self._plans.get(_lib.FLAVOR_DCONV, edge_index, edge_weight, num_nodes, flags=flags)

    def forward(self, X: torch.FloatTensor, edge_index: torch.LongTensor, edge_weight: torch.FloatTensor = None,
                cached_idx: bool = False) -> torch.FloatTensor:
        _require_cuda(X, "X")
        plan = self._plan(edge_index, edge_weight, X.size(-2))
        S = torch.cat(_basis(plan, X, self.weight.size(1)), dim=-1)
        H = torch.matmul(S, _stack_weight(self.weight))
        if self.bias is not None:
            H = H + self.bias
        return H


class BatchedDConv(DConv):
    """Reference: dcrnn.py:222-325 (degrees by scatter_add, duplicates legal).  `forward(X, edge_index,
    edge_weight, cached_idx)`: the plan cache subsumes `cached_idx`."""
    _batched_semantics = True


class DCRNN(torch.nn.Module):
    r"""Diffusion Convolutional GRU cell (reference: dcrnn.py:114-219).

    Args: in_channels, out_channels, K, bias -- as the reference (:128)."""

    _conv_cls = DConv
    _batched_semantics = False

    def __init__(self, in_channels: int, out_channels: int, K: int, bias: bool = True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.K = K
        self.bias = bias
        C = in_channels + out_channels
        self.conv_x_z = self._conv_cls(C, out_channels, K, bias)
        self.conv_x_r = self._conv_cls(C, out_channels, K, bias)
        self.conv_x_h = self._conv_cls(C, out_channels, K, bias)
        self._plans = PlanCache()
        self._wimg = ops.PackCache()

    # ---- helpers -----------------------------------------------------------------------------------
    def _plan(self, edge_index, edge_weight, num_nodes):
        flags = _lib.DCONV_ALLOW_DUPLICATES if self._batched_semantics else 0
        return self._plans.get(_lib.FLAVOR_DCONV, edge_index, edge_weight, num_nodes, flags=flags)

    def _needs_grad(self, *tensors):
        if not torch.is_grad_enabled():
            return False
        return any(p.requires_grad for p in self.parameters()) or any(t is not None and t.requires_grad for t in tensors)

    def _params(self):
        return (self.conv_x_z.weight, self.conv_x_r.weight, self.conv_x_h.weight,
                self.conv_x_z.bias, self.conv_x_r.bias, self.conv_x_h.bias)

    def _weight_image(self):
        """B-operand image for the tcgen05 kernel, rebuilt only when a parameter changes."""
        return self._wimg.get(list(self.parameters()),
                              lambda: ops.dcrnn_weight_image(*self._params(), self.in_channels, self.K))

    def _tiled_step(self, plan, X, H):
        """One GRU step on (N,*) or (B,N,*) tensors; z and r share the diffusion of [X|H]."""
        wz, wr, wh, bz, br, bh = self._params()
        O = self.out_channels
        S = torch.cat(_basis(plan, torch.cat([X, H], dim=-1), self.K), dim=-1)
        pre = torch.matmul(S, torch.cat([_stack_weight(wz), _stack_weight(wr)], dim=1))
        if bz is not None:
            pre = pre + torch.cat([bz, br])
        Z, R = torch.sigmoid(pre[..., :O]), torch.sigmoid(pre[..., O:])
        S2 = torch.cat(_basis(plan, torch.cat([X, H * R], dim=-1), self.K), dim=-1)
        ph = torch.matmul(S2, _stack_weight(wh))
        if bh is not None:
            ph = ph + bh
        Ht = torch.tanh(ph)
        return Z * H + (1 - Z) * Ht

    # ---- reference surface -------------------------------------------------------------------------
    def forward(self, X: torch.FloatTensor, edge_index: torch.LongTensor, edge_weight: torch.FloatTensor = None,
                H: torch.FloatTensor = None) -> torch.FloatTensor:
        """X (N,Cin), H (N,Cout) or None -> H' (N,Cout)   (dcrnn.py:194-219)."""
        _require_cuda(X, "X")
        N = X.shape[0]
        plan = self._plan(edge_index, edge_weight, N)
        if not self._needs_grad(X, H) and ops.dcrnn_seq_supported(plan, self.in_channels, self.out_channels, self.K):
            h0 = None if H is None else H.reshape(1, N, self.out_channels)
            out = ops.dcrnn_seq_fwd(plan, X.reshape(1, 1, N, self.in_channels), *self._params(), self.K, h0=h0,
                                    wimage=self._weight_image())
            return out[0, 0]
        if H is None:
            H = torch.zeros(N, self.out_channels, device=X.device, dtype=X.dtype)
        return self._tiled_step(plan, X, H)


class BatchedDCRNN(DCRNN):
    """Batched seq-to-seq DCRNN (reference: dcrnn.py:328-475).  X (B,T,N,Cin) -> (B,T,N,Cout), H_0 = 0.
    The reference replicates the graph B times block-diagonally (:363-369); the block-diagonal operator
    equals the single-graph operator applied per window, so nothing is replicated here."""

    _conv_cls = BatchedDConv
    _batched_semantics = True

    def forward(self, X, edge_index, edge_weight):
        _require_cuda(X, "X")
        B, T, N, F = X.size()
        plan = self._plan(edge_index, edge_weight, N)
        if not self._needs_grad(X) and ops.dcrnn_seq_supported(plan, self.in_channels, self.out_channels, self.K):
            try:
                return ops.dcrnn_seq_fwd(plan, X, *self._params(), self.K, wimage=self._weight_image())
            except _lib.StmpUnsupported:
                pass
        H = torch.zeros(B, N, self.out_channels, device=X.device, dtype=X.dtype)
        outs = []
        for t in range(T):
            H = self._tiled_step(plan, X[:, t], H)
            outs.append(H)
        return torch.stack(outs, dim=1)

    def forward_indexed(self, series, win_start, horizon, edge_index, edge_weight):
        """Index-batching entry: windows are read in-kernel from the resident series (T_total,N,Cin)
        at `win_start` (int64 [B]) -- the fused form of IndexDataset + DataLoader collate + forward."""
        _require_cuda(series, "series")
        plan = self._plan(edge_index, edge_weight, series.size(1))
        if not self._needs_grad(series) and ops.dcrnn_seq_supported(plan, self.in_channels, self.out_channels, self.K):
            return ops.dcrnn_seq_fwd(plan, series, *self._params(), self.K, win_start=win_start, horizon=horizon,
                                     wimage=self._weight_image())
        X = ops.window_gather(series, win_start, horizon, with_target=False)
        return self.forward(X, edge_index, edge_weight)
