"""Known-answer tests that pin the ORACLE itself: the COO gather/scatter restatement of the PyG
primitives must equal closed-form dense-matrix formulas (fp64), on hand-checkable graphs and on random
ones.  (PyG cannot be installed here, so this is the pin for the primitive semantics -- SURVEY 8c.)"""
import torch

from oracle import pyg, recurrent as R


def _graph(n=9, seed=0, loops=True):
    g = torch.Generator().manual_seed(seed)
    A = (torch.rand(n, n, generator=g) < 0.35).double() * (torch.rand(n, n, generator=g).double() * 0.9 + 0.1)
    A.fill_diagonal_(0)
    for i in range(n):
        A[i, (i + 1) % n] = 0.5
    if loops:
        A += torch.eye(n).double() * 0.7
    ei = A.nonzero().t().contiguous()
    return A, ei, A[ei[0], ei[1]]


def test_propagate_is_spmm_transpose():
    A, ei, w = _graph()
    x = torch.randn(9, 5, dtype=torch.double)
    # out[col] += w * x[row]  ==  A^T x
    assert torch.allclose(pyg.propagate(ei, x, w), A.t() @ x, atol=1e-12)
    xb = torch.randn(3, 9, 5, dtype=torch.double)
    assert torch.allclose(pyg.propagate(ei, xb, w), torch.einsum("rc,brf->bcf", A, xb), atol=1e-12)


def test_path_graph_by_hand():
    # 0 -> 1 -> 2 : propagate moves features one hop along the arrows
    ei = torch.tensor([[0, 1], [1, 2]])
    x = torch.tensor([[1.0], [10.0], [100.0]])
    out = pyg.propagate(ei, x, torch.tensor([2.0, 3.0]))
    assert out.flatten().tolist() == [0.0, 2.0, 30.0]


def test_to_dense_and_back():
    ei = torch.tensor([[0, 0, 2, 0], [1, 1, 0, 2]])
    adj = pyg.to_dense_adj(ei, edge_attr=torch.tensor([1.0, 2.0, 5.0, 7.0]))[0]
    assert adj.tolist() == [[0, 3, 7], [0, 0, 0], [5, 0, 0]]  # duplicates sum
    idx, val = pyg.dense_to_sparse(adj)
    assert idx.tolist() == [[0, 0, 2], [1, 2, 0]] and val.tolist() == [3, 7, 5]  # row-major


def test_get_laplacian_dense():
    A, ei, w = _graph(loops=True)
    n = A.size(0)
    A0 = A.clone().fill_diagonal_(0)
    deg = A0.sum(1)
    for norm, L in ((None, torch.diag(deg) - A0),
                    ("sym", torch.eye(n).double() - torch.diag(deg.pow(-0.5)) @ A0 @ torch.diag(deg.pow(-0.5))),
                    ("rw", torch.eye(n).double() - torch.diag(1 / deg) @ A0)):
        ei2, w2 = pyg.get_laplacian(ei, w, norm, num_nodes=n)
        dense = torch.zeros(n, n, dtype=torch.double).index_put_((ei2[0], ei2[1]), w2, accumulate=True)
        assert torch.allclose(dense, L, atol=1e-12), norm
        # loops are appended after the non-loop edges, in node order
        assert ei2[:, -n:].tolist() == [list(range(n)), list(range(n))]


def test_star_graph_sym_laplacian_by_hand():
    # the 4-node star of test/attention_test.py:186 (0-1, 0-2, 0-3, undirected)
    ei = torch.tensor([[0, 0, 0, 1, 2, 3], [1, 2, 3, 0, 0, 0]])
    ei2, w2 = pyg.get_laplacian(ei, None, "sym", torch.float32, 4)
    s = 1.0 / (3.0 ** 0.5)
    assert torch.allclose(w2, torch.tensor([-s] * 6 + [1.0] * 4))


def test_chebconv_dense():
    A, ei, w = _graph(loops=False)
    n = A.size(0)
    deg = A.sum(1)
    Lsym = torch.eye(n).double() - torch.diag(deg.pow(-0.5)) @ A @ torch.diag(deg.pow(-0.5))
    Lhat = (2.0 * Lsym / 2.0 - torch.eye(n).double())
    conv = pyg.ChebConv(3, 4, 4).double()
    x = torch.randn(n, 3, dtype=torch.double)
    # message flows row->col: out = Lhat^T x per hop
    M = Lhat.t()
    T = [x, M @ x]
    T.append(2 * M @ T[1] - T[0])
    T.append(2 * M @ T[2] - T[1])
    want = sum(T[k] @ conv.lins[k].weight.t() for k in range(4)) + conv.bias
    assert torch.allclose(conv(x, ei, w), want, atol=1e-10)


def test_gcnconv_dense():
    A, ei, w = _graph(loops=True)  # existing self loops keep their weight
    n = A.size(0)
    deg = A.sum(0)  # by column (target)
    Ahat = torch.diag(deg.pow(-0.5)) @ A @ torch.diag(deg.pow(-0.5))
    conv = pyg.GCNConv(3, 4).double()
    x = torch.randn(n, 3, dtype=torch.double)
    want = Ahat.t() @ (x @ conv.lin.weight.t()) + conv.bias
    assert torch.allclose(conv(x, ei, w), want, atol=1e-10)
    # missing loops are filled with 1 (2 if improved)
    A2, ei2, w2 = _graph(loops=False)
    for improved, fill in ((False, 1.0), (True, 2.0)):
        Af = A2 + torch.eye(n).double() * fill
        d = Af.sum(0)
        want = (torch.diag(d.pow(-0.5)) @ Af @ torch.diag(d.pow(-0.5))).t() @ (x @ conv.lin.weight.t()) + conv.bias
        c2 = pyg.GCNConv(3, 4, improved=improved).double()
        c2.load_state_dict(conv.state_dict())
        assert torch.allclose(c2(x, ei2, w2), want, atol=1e-10)


def test_dconv_dense_symmetric_pattern():
    """On a pattern-symmetric, row-major edge list the positional norm_in pairing is self-consistent and
    DConv equals D_out^-1/D_in^-1 random-walk matrices built from the 0/1 pattern."""
    n = 8
    g = torch.Generator().manual_seed(1)
    P = (torch.rand(n, n, generator=g) < 0.4)
    P = (P | P.t()) | torch.eye(n, dtype=torch.bool)
    W = torch.rand(n, n, generator=g).double() * P
    ei = P.nonzero().t().contiguous()
    w = W[ei[0], ei[1]].float()
    ops = R.dconv_operators(ei, w, batched=False, num_nodes=n)
    x = torch.randn(n, 3)
    To = pyg.propagate(ops[0], x, ops[1])
    Ti = pyg.propagate(ops[2], x, ops[3])
    deg_out, deg_in = W.sum(1).float(), W.sum(0).float()
    Pf = P.float()
    # To[c] = sum_r P[r,c] x[r]/deg_out[r];  Ti[r] = sum_c P[r,c] x[c]/deg_in[c]
    assert torch.allclose(To, Pf.t() @ (x / deg_out[:, None]), atol=1e-5)
    assert torch.allclose(Ti, Pf @ (x / deg_in[:, None]), atol=1e-5)


def test_dconv_zero_degree_gives_inf():
    ei = torch.tensor([[0, 1], [1, 2]])  # node 2 has no outgoing edge, node 0 no incoming
    ops = R.dconv_operators(ei, None, batched=False, num_nodes=3)
    assert torch.isfinite(ops[1]).all()          # norm_out = 1/deg_out[row]: rows 0,1 have out-degree 1
    assert torch.isinf(ops[3]).any()             # norm_in = 1/deg_in[row]: row 0 has in-degree 0 -> inf
