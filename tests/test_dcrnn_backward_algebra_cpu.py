"""Host-side algebra of the hand-written DCRNN backward (`nn/recurrent/dcrnn.py::_DcrnnSeqFn.backward`, per-step variant)
checked on the CPU against autograd: the CUDA entry points it calls are replaced by dense torch stand-ins that follow the
documented contracts of include/stmp.h (stmp_spmm with column blocks, stmp_gru_bwd_carry / _zr, stmp_dcrnn_pack_bwd_weights).
What this pins without a GPU: the reverse-time recurrence, the in-place basis adjoint, the hoisted weight-gradient GEMMs and
the un-stacking of the stacked-basis gradients, for K = 1..4 with and without an initial state."""
import pytest
import torch

import pytorch_geometric_temporal_b200.nn.recurrent.dcrnn as D


class _DenseOps(object):
    """Dense stand-ins: P[0], P[1] are the two diffusion operators as (N,N) matrices."""

    def __init__(self, P):
        self.P = P

    def spmm_raw(self, plan, op, x, transposed=False, alpha=1.0, z=None, beta=0.0, att=None, out=None):
        A = self.P[op].t() if transposed else self.P[op]
        y = alpha * torch.matmul(A, x)
        return y if z is None else y + beta * z

    def spmm(self, plan, op, x, alpha=1.0, z=None, beta=0.0, att=None):
        return self.spmm_raw(plan, op, x, False, alpha, z, beta)

    def spmm_cols(self, plan, op, buf, src, dst, width, alpha=1.0, z_col=None, beta=0.0, transposed=False):
        A = self.P[op].t() if transposed else self.P[op]
        y = alpha * torch.matmul(A, buf[..., src:src + width])
        if z_col is not None:
            y = y + beta * buf[..., z_col:z_col + width]
        buf[..., dst:dst + width] = y

    def dcrnn_bwd_supported(self, plan, cin, cout, K):
        return False                                            # the persistent kernel path needs CUDA streams

    def dcrnn_pack_bwd_weights(self, wz, wr, wh, cin, K):
        return D._stack_weight(wh).t().contiguous(), torch.cat([D._stack_weight(wz), D._stack_weight(wr)], dim=1).t().contiguous()

    def gru_bwd_carry(self, cin, cout, du2, du1, g_prev=None, z_prev=None, r_prev=None, dx=None, gout=None, z=None, ht=None, g=None,
                      dph=None, dh_out=None):
        dh = 0
        if g_prev is not None:
            dh = g_prev * z_prev + du2[..., cin:cin + cout] * r_prev + du1[..., cin:cin + cout]
            if dx is not None:
                dx.copy_(du2[..., :cin] + du1[..., :cin])
        if dh_out is not None:
            dh_out.copy_(dh)
        if gout is not None:
            gg = gout + dh
            g.copy_(gg)
            dph.copy_(gg * (1 - z) * (1 - ht * ht))

    def gru_bwd_zr(self, cin, cout, g, hprev, z, r, ht, du2, dpzr):
        hp = hprev if hprev is not None else 0
        dpzr[..., :cout] = g * (hp - ht) * z * (1 - z)
        dpzr[..., cout:] = du2[..., cin:cin + cout] * hp * r * (1 - r)


class _Ctx(object):
    pass


@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("use_h0", [True, False])
def test_per_step_backward_matches_autograd(monkeypatch, K, use_h0):
    torch.manual_seed(K)
    N, Ci, Co, B, T = 7, 2, 5, 2, 4
    P = [torch.rand(N, N) * (torch.rand(N, N) < 0.5) for _ in range(2)]
    monkeypatch.setattr(D, "ops", _DenseOps(P))
    wz, wr, wh = [(torch.randn(2, K, Ci + Co, Co) * 0.3).requires_grad_(True) for _ in range(3)]
    bz, br, bh = [torch.randn(Co, requires_grad=True) for _ in range(3)]
    X = torch.randn(B, T, N, Ci, requires_grad=True)
    H0 = torch.randn(B, N, Co, requires_grad=True)
    # forward in plain torch (the cell of dcrnn.py:172-219 on the stacked basis), recording the gate stash
    H, outs, st = (H0 if use_h0 else torch.zeros(B, N, Co)), [], []
    for t in range(T):
        S = torch.cat(D._basis(None, torch.cat([X[:, t], H], -1), K), -1)
        pre = S @ torch.cat([D._stack_weight(wz), D._stack_weight(wr)], 1) + torch.cat([bz, br])
        Z, R = torch.sigmoid(pre[..., :Co]), torch.sigmoid(pre[..., Co:])
        S2 = torch.cat(D._basis(None, torch.cat([X[:, t], H * R], -1), K), -1)
        Ht = torch.tanh(S2 @ D._stack_weight(wh) + bh)
        H = Z * H + (1 - Z) * Ht
        outs.append(H)
        st.append(torch.stack([Z, R, Ht], 1))
    out, stash = torch.stack(outs, 1), torch.stack(st, 1)
    gout = torch.randn_like(out)
    out.backward(gout)
    leaves = (X, H0, wz, wr, wh, bz, br, bh)
    want = [t.grad.clone() if t.grad is not None else None for t in leaves]
    ctx = _Ctx()
    ctx.plan, ctx.K, ctx.has_bias, ctx.has_h0, ctx.needs_input_grad = None, K, True, use_h0, [True, True]
    ctx.saved_tensors = (X.detach(), H0.detach() if use_h0 else None, wz.detach(), wr.detach(), wh.detach(), out.detach(), stash.detach())
    with torch.no_grad():
        got = D._DcrnnSeqFn.backward(ctx, gout)
    for name, w, g in zip(("X", "H0", "wz", "wr", "wh", "bz", "br", "bh"), want, got[:8]):
        if w is None:
            assert g is None, name
        else:
            assert torch.allclose(g, w, rtol=1e-4, atol=1e-5), (name, float((g - w).abs().max()))
