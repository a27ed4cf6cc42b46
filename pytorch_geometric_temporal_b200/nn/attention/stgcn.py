"""STConv / TemporalConv -- drop-in for torch_geometric_temporal/nn/attention/stgcn.py (TemporalConv :8-47, STConv
:50-160).  Same constructors, forward signatures and state_dict keys
(`_temporal_conv{1,2}.conv_{1,2,3}.{weight,bias}`, `_graph_conv.lins.{k}.weight`, `_graph_conv.bias`,
`_batch_norm.*`).

The reference calls ChebConv once per (batch row, timestep) slice in a Python double loop (:151-153), renormalising
the Laplacian every time.  Here the (B, T') slices are the batch axis of ONE `stmp_spmm` launch per Chebyshev hop on
the cached plan, and the K products are one GEMM.  The gated (1,k) convolutions are fp32 contractions
(`_conv_1xk`: cuDNN would silently use TF32 and miss the 1e-4 parity bar)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..recurrent._cheb import ChebConv
from .astgcn import _conv_1xk


class TemporalConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3):
        super().__init__()
        self.conv_1 = nn.Conv2d(in_channels, out_channels, (1, kernel_size))
        self.conv_2 = nn.Conv2d(in_channels, out_channels, (1, kernel_size))
        self.conv_3 = nn.Conv2d(in_channels, out_channels, (1, kernel_size))

    def forward(self, X: torch.FloatTensor) -> torch.FloatTensor:
        """X (B, T, N, C_in) -> (B, T-k+1, N, C_out): relu(P * sigmoid(Q) + R)   (stgcn.py:28-47)."""
        Xp = X.permute(0, 3, 2, 1)
        P = _conv_1xk(self.conv_1, Xp)
        Q = torch.sigmoid(_conv_1xk(self.conv_2, Xp))
        H = F.relu(P * Q + _conv_1xk(self.conv_3, Xp))
        return H.permute(0, 3, 2, 1)


class STConv(nn.Module):
    def __init__(self, num_nodes: int, in_channels: int, hidden_channels: int, out_channels: int, kernel_size: int, K: int,
                 normalization: str = "sym", bias: bool = True):
        super().__init__()
        self.num_nodes, self.in_channels, self.hidden_channels, self.out_channels = num_nodes, in_channels, hidden_channels, out_channels
        self.kernel_size, self.K, self.normalization, self.bias = kernel_size, K, normalization, bias
        self._temporal_conv1 = TemporalConv(in_channels, hidden_channels, kernel_size)
        self._graph_conv = ChebConv(hidden_channels, hidden_channels, K, normalization=normalization, bias=bias)
        self._temporal_conv2 = TemporalConv(hidden_channels, out_channels, kernel_size)
        self._batch_norm = nn.BatchNorm2d(num_nodes)

    def forward(self, X: torch.FloatTensor, edge_index: torch.LongTensor, edge_weight: torch.FloatTensor = None) -> torch.FloatTensor:
        T_0 = self._temporal_conv1(X)                                      # (B, T', N, hidden)
        B, Tp, N, Ch = T_0.shape
        T = self._graph_conv(T_0.reshape(B * Tp, N, Ch), edge_index, edge_weight).reshape(B, Tp, N, Ch)
        T = self._temporal_conv2(F.relu(T))
        T = self._batch_norm(T.permute(0, 2, 1, 3))
        return T.permute(0, 2, 1, 3)
