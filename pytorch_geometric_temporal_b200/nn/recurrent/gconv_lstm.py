"""GConvLSTM -- drop-in for torch_geometric_temporal/nn/recurrent/gconv_lstm.py (:9-238): constructor
`(in_channels, out_channels, K, normalization="sym", bias=True)`, `forward(X, edge_index, edge_weight,
H, C, lambda_max) -> (H, C)`, state_dict keys `conv_{x,h}_{i,f,c,o}.lins.{k}.weight/.bias`,
`w_c_{i,f,o} (1,out)` (glorot), `b_{i,f,c,o} (1,out)` (zeros).  Eight ChebConvs per step in the
reference = 8(K-1) propagations; here T_k([X|H]) is computed once (K-1 SpMMs on Ci+Co channels) and one
GEMM produces all four gate pre-activations."""
import torch

from ... import _lib, ops
from ...plan import _require_cuda
from ._cheb import ChebParams, ChebPlanMixin, cheb_basis, glorot_


class GConvLSTM(torch.nn.Module, ChebPlanMixin):
    def __init__(self, in_channels: int, out_channels: int, K: int, normalization: str = "sym", bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels, self.K = in_channels, out_channels, K
        self.normalization, self.bias = normalization, bias
        P = torch.nn.Parameter
        # creation order mirrors the reference (gconv_lstm.py:60-147) so seeded init consumes the RNG identically
        for g in "ifco":
            setattr(self, f"conv_x_{g}", ChebParams(in_channels, out_channels, K, bias))
            setattr(self, f"conv_h_{g}", ChebParams(out_channels, out_channels, K, bias))
            if g != "c":
                setattr(self, f"w_c_{g}", P(torch.empty(1, out_channels)))
            setattr(self, f"b_{g}", P(torch.empty(1, out_channels)))
        for g in "ifo":
            glorot_(getattr(self, f"w_c_{g}"))
        for g in "ifco":
            torch.nn.init.zeros_(getattr(self, f"b_{g}"))
        self._init_plans()
        self._pack = ops.PackCache()

    def _weight(self):
        cols = []
        for g in "ifco":
            cx, ch = getattr(self, f"conv_x_{g}"), getattr(self, f"conv_h_{g}")
            cols.append(torch.cat([torch.cat([cx.lins[k].weight.t(), ch.lins[k].weight.t()], dim=0) for k in range(self.K)], dim=0))
        return torch.cat(cols, dim=1)

    def _conv_bias(self):
        if self.conv_x_i.bias is None:
            return None
        return torch.cat([getattr(self, f"conv_x_{g}").bias + getattr(self, f"conv_h_{g}").bias for g in "ifco"])

    def forward(self, X: torch.FloatTensor, edge_index: torch.LongTensor, edge_weight: torch.FloatTensor = None,
                H: torch.FloatTensor = None, C: torch.FloatTensor = None, lambda_max: torch.Tensor = None):
        _require_cuda(X, "X")
        N, Co = X.size(-2), self.out_channels
        if H is None:
            H = torch.zeros(*X.shape[:-1], Co, device=X.device, dtype=X.dtype)
        if C is None:
            C = torch.zeros(*X.shape[:-1], Co, device=X.device, dtype=X.dtype)
        plan = self._cheb_plan(edge_index, edge_weight, N, self.normalization, lambda_max)
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or X.requires_grad
                                                  or H.requires_grad or C.requires_grad)
        Cw = self.in_channels + Co
        if not needs_grad and Co in (32, 64) and (self.K * Cw) % 4 == 0 and Cw % 4 == 0:
            # large-graph inference: T_k written in place into S = [T_0|T_1|..] by the SpMM kernel, then ONE tcgen05
            # launch does S @ W and the whole peephole-LSTM gate chain in its epilogue (stmp_gemm_lstm_f32)
            S = torch.empty(*X.shape[:-1], self.K * Cw, device=X.device, dtype=torch.float32)
            S[..., :self.in_channels] = X
            S[..., self.in_channels:Cw] = H
            for k in range(1, self.K):
                if k == 1:
                    ops.spmm_cols(plan, 0, S, 0, Cw, Cw)
                else:
                    ops.spmm_cols(plan, 0, S, (k - 1) * Cw, k * Cw, Cw, alpha=2.0, z_col=(k - 2) * Cw, beta=-1.0)
            packed, cb = self._pack.get(list(self.parameters()), lambda: (ops.gemm_prepack(self._weight()), self._conv_bias()))
            try:
                return ops.gemm_lstm(S, packed, self.K * Cw, Co, cb, C.contiguous(), self.w_c_i, self.w_c_f, self.w_c_o,
                                     self.b_i, self.b_f, self.b_c, self.b_o)
            except _lib.StmpUnsupported:
                pass
        S = torch.cat(cheb_basis(plan, torch.cat([X, H], dim=-1), self.K), dim=-1)
        pre = torch.matmul(S, self._weight())
        cb = self._conv_bias()
        if cb is not None:
            pre = pre + cb
        pi, pf, pc, po = (pre[..., j * Co:(j + 1) * Co] for j in range(4))
        grad = needs_grad
        if grad:
            I = torch.sigmoid(pi + self.w_c_i * C + self.b_i)
            Fg = torch.sigmoid(pf + self.w_c_f * C + self.b_f)
            Cn = Fg * C + I * torch.tanh(pc + self.b_c)
            O = torch.sigmoid(po + self.w_c_o * Cn + self.b_o)          # uses the NEW cell state (:235-236)
            return O * torch.tanh(Cn), Cn
        Cn = ops.lstm_ifc(pi.contiguous(), pf.contiguous(), pc.contiguous(), C, self.w_c_i, self.w_c_f, self.b_i, self.b_f, self.b_c)
        Hn = ops.lstm_oh(po.contiguous(), Cn, self.w_c_o, self.b_o)
        return Hn, Cn
