"""MSTGCN -- drop-in for torch_geometric_temporal/nn/attention/mstgcn.py (MSTGCNBlock :10-122, MSTGCN :125-200;
SURVEY 8f rank 1).  Same constructors, forward signatures and state_dict keys (`_blocklist.{i}._cheb_conv.lins.{k}.
weight/.bias`, `._time_conv.*`, `._residual_conv.*`, `._layer_norm.*`, `_final_conv.*`).

Graph work goes through one batched `stmp_spmm` per Chebyshev hop on the cached plan (all B*T slices at once);
lambda_max -- scipy ARPACK on the host in EVERY block forward in the reference (:79-81) -- is computed once per
static graph.  The reference's (F,B,N,T) -> (N,F,T*B) `reshape` (:84-88) reinterprets memory rather than transposing;
that is the observable behaviour and it is reproduced literally."""
from typing import List, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..recurrent._cheb import ChebConv
from .astgcn import _conv_1xk, _reset, laplacian_lambda_max


class MSTGCNBlock(nn.Module):
    def __init__(self, in_channels: int, K: int, nb_chev_filter: int, nb_time_filter: int, time_strides: int):
        super().__init__()
        self._cheb_conv = ChebConv(in_channels, nb_chev_filter, K, normalization=None)
        self._time_conv = nn.Conv2d(nb_chev_filter, nb_time_filter, kernel_size=(1, 3), stride=(1, time_strides), padding=(0, 1))
        self._residual_conv = nn.Conv2d(in_channels, nb_time_filter, kernel_size=(1, 1), stride=(1, time_strides))
        self._layer_norm = nn.LayerNorm(nb_time_filter)
        self.nb_time_filter = nb_time_filter
        self._lam_cache = {}
        _reset(self)

    def _lambda_max(self, edge_index, num_nodes):
        key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape))
        hit = self._lam_cache.get(key)
        if hit is None:
            if len(self._lam_cache) > 16:
                self._lam_cache.clear()
            hit = (laplacian_lambda_max(edge_index, num_nodes, None), edge_index)   # LaplacianLambdaMax() default: L = D - A
            self._lam_cache[key] = hit
        return hit[0]

    def forward(self, X: torch.FloatTensor, edge_index: Union[torch.LongTensor, List[torch.LongTensor]]) -> torch.FloatTensor:
        B, N, Fi, T = X.shape
        if not isinstance(edge_index, list):
            lam = self._lambda_max(edge_index, N)
            Xt = X.permute(2, 0, 1, 3).reshape(N, Fi, T * B).permute(2, 0, 1)        # the reference's reinterpretation
            Xt = F.relu(self._cheb_conv(Xt.contiguous(), edge_index, lambda_max=lam))   # (T*B, N, Fc): one batched pass
            Xt = Xt.permute(1, 2, 0).reshape(N, self.nb_time_filter, B, T).permute(2, 0, 1, 3)
        else:
            hats = []
            for t in range(T):
                lam = self._lambda_max(edge_index[t], N)
                hats.append(self._cheb_conv(X[:, :, :, t].contiguous(), edge_index[t], lambda_max=lam).unsqueeze(-1))
            Xt = F.relu(torch.cat(hats, dim=-1))
        Xt = _conv_1xk(self._time_conv, Xt.permute(0, 2, 1, 3))
        Xr = _conv_1xk(self._residual_conv, X.permute(0, 2, 1, 3))
        Y = self._layer_norm(F.relu(Xr + Xt).permute(0, 3, 2, 1))
        return Y.permute(0, 2, 3, 1)


class MSTGCN(nn.Module):
    def __init__(self, nb_block: int, in_channels: int, K: int, nb_chev_filter: int, nb_time_filter: int, time_strides: int,
                 num_for_predict: int, len_input: int):
        super().__init__()
        self._blocklist = nn.ModuleList([MSTGCNBlock(in_channels, K, nb_chev_filter, nb_time_filter, time_strides)])
        self._blocklist.extend([MSTGCNBlock(nb_time_filter, K, nb_chev_filter, nb_time_filter, 1) for _ in range(nb_block - 1)])
        self._final_conv = nn.Conv2d(int(len_input / time_strides), num_for_predict, kernel_size=(1, nb_time_filter))
        _reset(self)

    def forward(self, X: torch.FloatTensor, edge_index: torch.LongTensor) -> torch.FloatTensor:
        """X (B, N, F_in, T_in) -> (B, N, T_out)   (mstgcn.py:181-200)."""
        for block in self._blocklist:
            X = block(X, edge_index)
        Y = _conv_1xk(self._final_conv, X.permute(0, 3, 1, 2))            # kernel (1, F_t): contraction over (T_in, F_t)
        return Y[:, :, :, -1].permute(0, 2, 1)

