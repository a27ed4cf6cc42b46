"""GConvGRU -- drop-in for torch_geometric_temporal/nn/recurrent/gconv_gru.py (:5-170): same
constructor `(in_channels, out_channels, K, normalization="sym", bias=True)`, `forward(X, edge_index,
edge_weight=None, H=None, lambda_max=None)`, state_dict keys `conv_{x,h}_{z,r,h}.lins.{k}.weight`,
`.bias`.  The six ChebConvs of the reference renormalise the graph and re-propagate per gate; here the
scaled Laplacian is a cached plan and T_k([X|H]) is computed once and shared by all gates."""
import torch

from ... import ops
from ...plan import _require_cuda
from ._cheb import ChebParams, ChebPlanMixin, cheb_basis


class GConvGRU(torch.nn.Module, ChebPlanMixin):
    def __init__(self, in_channels: int, out_channels: int, K: int, normalization: str = "sym", bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels, self.K = in_channels, out_channels, K
        self.normalization, self.bias = normalization, bias
        for g in "zrh":
            setattr(self, f"conv_x_{g}", ChebParams(in_channels, out_channels, K, bias))
            setattr(self, f"conv_h_{g}", ChebParams(out_channels, out_channels, K, bias))
        self._init_plans()
        self._pack = ops.PackCache()

    def _gate_weight(self, g, x_only=False, h_only=False):
        """Rows follow the basis [T_0 | T_1 | ...] of U=[X|H]: block k = [Wx_k^T ; Wh_k^T]."""
        cx, ch = getattr(self, f"conv_x_{g}"), getattr(self, f"conv_h_{g}")
        rows = []
        for k in range(self.K):
            if not h_only:
                rows.append(cx.lins[k].weight.t())
            if not x_only:
                rows.append(ch.lins[k].weight.t())
        return torch.cat(rows, dim=0)

    def _gate_bias(self, g):
        cx, ch = getattr(self, f"conv_x_{g}"), getattr(self, f"conv_h_{g}")
        return None if cx.bias is None else cx.bias + ch.bias

    def _packed(self):
        """(wcat [96,112], bcat [96]) for stmp_gru_seq_fwd: columns H | L^H | - | X | L^X | - (one operator)."""
        def build():
            Ci, dev = self.in_channels, self.conv_x_z.lins[0].weight.device
            W = torch.zeros(96, 112, device=dev)
            b = torch.zeros(96, device=dev)
            for gi, g in enumerate("zrh"):
                cx, ch = getattr(self, f"conv_x_{g}"), getattr(self, f"conv_h_{g}")
                r = slice(32 * gi, 32 * gi + 32)
                W[r, 0:32] = ch.lins[0].weight
                W[r, 96:96 + Ci] = cx.lins[0].weight
                if self.K > 1:
                    W[r, 32:64] = ch.lins[1].weight
                    W[r, 100:100 + Ci] = cx.lins[1].weight
                if cx.bias is not None:
                    b[r] = cx.bias + ch.bias
            return W, b, ops.gru_weight_image(W, b)
        return self._pack.get(list(self.parameters()), build)

    def _fused_ok(self, plan, X, H):
        if self.K > 2 or self.out_channels != 32 or X.dim() != 2:
            return False
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or X.requires_grad
                                        or (H is not None and H.requires_grad)):
            return False
        return ops.gru_seq_supported(plan, 1 if self.K > 1 else 0, self.in_channels, self.out_channels)

    def forward(self, X: torch.FloatTensor, edge_index: torch.LongTensor, edge_weight: torch.FloatTensor = None,
                H: torch.FloatTensor = None, lambda_max: torch.Tensor = None) -> torch.FloatTensor:
        _require_cuda(X, "X")
        N, Ci, Co, K = X.size(-2), self.in_channels, self.out_channels, self.K
        if H is None:
            H = torch.zeros(*X.shape[:-1], Co, device=X.device, dtype=X.dtype)
        plan = self._cheb_plan(edge_index, edge_weight, N, self.normalization, lambda_max)
        if self._fused_ok(plan, X, H):   # one tcgen05 launch for the whole cell (stmp_gru_seq_fwd)
            W, b, img = self._packed()
            return ops.gru_seq_fwd(plan, 1 if K > 1 else 0, X.reshape(1, 1, N, Ci), W, b, h0=H.reshape(1, N, Co), wimage=img)[0, 0]
        TU = cheb_basis(plan, torch.cat([X, H], dim=-1), K)              # K x (N, Ci+Co)
        S = torch.cat(TU, dim=-1)
        pre = torch.matmul(S, torch.cat([self._gate_weight("z"), self._gate_weight("r")], dim=1))
        bz, br = self._gate_bias("z"), self._gate_bias("r")
        if bz is not None:
            pre = pre + torch.cat([bz, br])
        grad = torch.is_grad_enabled() and (pre.requires_grad or H.requires_grad)
        if grad:
            Z, R = torch.sigmoid(pre[..., :Co]), torch.sigmoid(pre[..., Co:])
            HR = H * R
        else:
            Z, R, HR = ops.gru_zr(pre[..., :Co].contiguous(), pre[..., Co:].contiguous(), H)
        Sx = torch.cat([t[..., :Ci] for t in TU], dim=-1)                   # T_k(X) is shared with the candidate
        Shr = torch.cat(cheb_basis(plan, HR, K), dim=-1)
        ph = torch.matmul(Sx, self._gate_weight("h", x_only=True)) + torch.matmul(Shr, self._gate_weight("h", h_only=True))
        bh = self._gate_bias("h")
        if bh is not None:
            ph = ph + bh
        if grad:
            return Z * H + (1 - Z) * torch.tanh(ph)
        return ops.gru_out(ph, Z, H)
