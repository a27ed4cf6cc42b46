"""Round-2 golden vectors from the UNMODIFIED reference modules (same mechanism as make_goldens.py: /root/reference imported
through oracle/refload.py on top of oracle/stubs).  Run in the build container only:  python tests/golden/make_goldens_r2.py [name ...]

* dcrnn_cfg2_grads   -- BatchedDCRNN(2,32,K=2) at the METR-LA shape: output AND autograd gradients (input + parameters)
* a3tgcn2_cfg3       -- A3TGCN2(2,32,12,B) at the PEMS-BAY shape (325 nodes), all rows, with and without an incoming H
* a3tgcn2_cfg3_grads -- the same shape without an incoming state: outputs and autograd gradients of every parameter
* astgcn_cfg4        -- ASTGCN(3 blocks, K=3, 64/64 filters) at the PeMS04 shape (307 nodes, B=32 rows subsampled to keep the
                        file small: the model is row-independent, so B rows of the reference output are exact for those rows)
* gconv_lstm_cfg5seq -- GConvLSTM(64,64,K=3) 12-step recurrence on a 2 000-node slice-shaped graph (CPU-tractable) + grads
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refload, pyg  # noqa: E402
from pytorch_geometric_temporal_b200.dataset import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **kw):
    torch.save(kw, os.path.join(OUT, name + ".pt"))
    print(f"{name}.pt  {os.path.getsize(os.path.join(OUT, name + '.pt')) / 1024:.0f} KB")


def sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def dcrnn_cfg2_grads():
    dc = refload.load("nn.recurrent.dcrnn")
    ei, ew, series = synthetic.metr_la_like(seed=0, t_total=64)
    ei_t, ew_t = torch.from_numpy(ei), torch.from_numpy(ew)
    torch.manual_seed(0)
    m = dc.BatchedDCRNN(2, 32, 2)
    X = torch.from_numpy(np.stack([series[3:15], series[20:32], series[41:53]])).clone().requires_grad_(True)  # (3,12,207,2)
    out = m(X, ei_t, ew_t)
    w = torch.linspace(-1, 1, out.numel()).view_as(out)
    (out * w).sum().backward()
    save("dcrnn_cfg2_grads", edge_index=ei_t, edge_weight=ew_t, X=X.detach(), state=sd(m), out=out.detach(), gX=X.grad.clone(),
         grads={k: p.grad.detach().clone() for k, p in m.named_parameters()}, K=2)


def a3tgcn2_cfg3():
    at = refload.load("nn.recurrent.attentiontemporalgcn")
    tg = refload.load("nn.recurrent.temporalgcn")
    ei, ew, _ = synthetic.pems_bay_like(0, 16)
    ei_t, ew_t = torch.from_numpy(ei), torch.from_numpy(ew)
    torch.manual_seed(0)
    m = at.A3TGCN2(2, 32, 12, 64)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(8, 325, 2, 12, generator=g)
    H = torch.randn(8, 325, 32, generator=g) * 0.5
    m1 = at.A3TGCN(2, 32, 12)
    c2 = tg.TGCN2(2, 32, 8)
    with torch.no_grad():
        out, outH = m(X, ei_t, ew_t), m(X, ei_t, ew_t, H)
        out1, out1H = m1(X[0], ei_t, ew_t), m1(X[0], ei_t, ew_t, H[0])
        cell, cellH = c2(X[..., 0], ei_t, ew_t), c2(X[..., 0], ei_t, ew_t, H)
    save("a3tgcn2_cfg3", edge_index=ei_t, edge_weight=ew_t, X=X, H=H, state=sd(m), out=out, outH=outH,
         state1=sd(m1), out1=out1, out1H=out1H, state_cell=sd(c2), cell=cell, cellH=cellH)


def a3tgcn2_cfg3_grads():
    """A3TGCN2 / TGCN2 at the PEMS-BAY shape WITHOUT an incoming state (the training call of the reference's A3TGCN2 example): outputs and
    autograd gradients of every parameter of the unmodified reference."""
    at = refload.load("nn.recurrent.attentiontemporalgcn")
    tg = refload.load("nn.recurrent.temporalgcn")
    ei, ew, _ = synthetic.pems_bay_like(0, 16)
    ei_t, ew_t = torch.from_numpy(ei), torch.from_numpy(ew)
    torch.manual_seed(3)
    m = at.A3TGCN2(2, 32, 12, 8)
    g = torch.Generator().manual_seed(9)
    X = torch.randn(8, 325, 2, 12, generator=g)
    out = m(X, ei_t, ew_t)
    w = torch.linspace(-1, 1, out.numel()).view_as(out)
    (out * w).sum().backward()
    c2 = tg.TGCN2(2, 32, 8)
    cell = c2(X[..., 3], ei_t, ew_t)
    (cell * w).sum().backward()
    save("a3tgcn2_cfg3_grads", edge_index=ei_t, edge_weight=ew_t, X=X, state=sd(m), out=out.detach(),
         grads={k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None},
         state_cell=sd(c2), cell=cell.detach(), grads_cell={k: p.grad.detach().clone() for k, p in c2.named_parameters() if p.grad is not None})


def astgcn_cfg4():
    """BASELINE configs[3]: ASTGCN(3 blocks, K=3, 64 Chebyshev / 64 time filters, stride 1, 12 -> 12) on the PeMS04 shape (307 nodes,
    340 undirected links), batch 32, normalization "sym"; plus normalization None (lambda_max by scipy) on 8 rows."""
    mod = refload.load("nn.attention.astgcn")
    eiu = torch.from_numpy(synthetic.pems04_like(0))
    cases = {}
    for norm, B in (("sym", 32), (None, 8)):
        torch.manual_seed(0)
        m = mod.ASTGCN(3, 1, 3, 64, 64, 1, 12, 12, 307, normalization=norm)
        g = torch.Generator().manual_seed(11)
        X = torch.randn(B, 307, 1, 12, generator=g)
        with torch.no_grad():
            out = m(X, eiu)
        # Vs / bs are 3 x 2 x 307 x 307 floats: regenerate them from the seed at test time instead of storing them
        cases[str(norm)] = dict(normalization=norm, X=X, out=out, seed=0)
        small = {k: v.detach().clone() for k, v in m.state_dict().items()}
        cases[str(norm)]["state_checksum"] = float(sum(v.double().abs().sum() for v in small.values()))
    save("astgcn_cfg4", edge_index=eiu, cases=cases,
         ctor=dict(nb_block=3, in_channels=1, K=3, nb_chev_filter=64, nb_time_filter=64, time_strides=1, num_for_predict=12,
                   len_input=12, num_of_vertices=307))


def gconv_lstm_cfg5seq():
    """GConvLSTM(64,64,K=3) (the cfg5 cell) unrolled over 6 steps on a 600-node / 6000-edge random graph (CPU-tractable slice of the
    10k/100k shape): final H, C and the autograd gradients of every parameter and of X from the UNMODIFIED reference module."""
    gl = refload.load("nn.recurrent.gconv_lstm")
    ei, ew = synthetic.large_graph(600, 6000, 1)
    ei_t, ew_t = torch.from_numpy(ei), torch.from_numpy(ew)
    torch.manual_seed(3)
    m = gl.GConvLSTM(64, 64, 3)
    for n_, p in m.named_parameters():                 # non-zero biases / peepholes so their gradients are exercised
        if n_.startswith("b_") or n_.endswith(".bias"):
            torch.nn.init.normal_(p, std=0.1)
    g = torch.Generator().manual_seed(9)
    X = (torch.randn(6, 600, 64, generator=g) * 0.5).requires_grad_(True)
    H = C = None
    loss = 0
    for t in range(6):
        H, C = m(X[t], ei_t, ew_t, H, C)
        loss = loss + (H * torch.linspace(-1, 1, H.numel()).view_as(H)).sum() + 0.3 * C.square().sum()
    loss.backward()
    save("gconv_lstm_cfg5seq", edge_index=ei_t, edge_weight=ew_t, X=X.detach(), state=sd(m), H=H.detach(), C=C.detach(),
         gX=X.grad.clone(), grads={k: p.grad.detach().clone() for k, p in m.named_parameters()})


GENERATORS = {"dcrnn_cfg2_grads": dcrnn_cfg2_grads, "a3tgcn2_cfg3": a3tgcn2_cfg3, "a3tgcn2_cfg3_grads": a3tgcn2_cfg3_grads, "astgcn_cfg4": astgcn_cfg4,
              "gconv_lstm_cfg5seq": gconv_lstm_cfg5seq}


if __name__ == "__main__":
    names = sys.argv[1:] or list(GENERATORS)
    for n in names:
        GENERATORS[n]()
