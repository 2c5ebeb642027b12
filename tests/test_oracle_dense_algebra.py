"""CPU tests: the oracle's three convs against an INDEPENDENT dense-matrix statement of the published layer equations.

`oracle/pyg_restated.py` restates PyG 2.5.3's gather / scatter code path (the reference's convs,
`/root/reference/src/models/gnn.py:20-23,41-44,64-67`; PyG itself is not installable here).  These tests evaluate the
same layers a second way that shares no code with it -- dense [N, N] adjacency algebra in float64, straight from the
layer definitions --

  GCN   (Kipf & Welling, eq. 2):   H' = D~^-1/2 (A + I) D~^-1/2 X W + b,      D~ = rowsum(A + I)
  SAGE  (Hamilton et al., mean):   H' = (D^-1 A X) W_l + b_l + X W_r          (rows without in-edges aggregate to 0)
  GAT   (Velickovic et al., 1-4):  e_ij = LeakyReLU(a_dst.Wx_i + a_src.Wx_j),  alpha = softmax_j over N(i) + {i}

with the multigraph conventions PyG's code implies (and SURVEY.md Appendix A states): A[i, j] counts the parallel
edges j -> i; existing self loops are dropped and exactly one of weight 1 per node is added (`add_remaining_self_loops`
for GCN, `remove_self_loops` + `add_self_loops` for GAT); a duplicate edge is its own softmax entry.  Forward values and
the gradients of a random scalar functional are compared, so the scatter path's normalisation, self-loop rewrite,
mean denominator, softmax stabilisation (+1e-16, detached max) and bias placement are all pinned from a second side.
"""
import pytest
import torch

from oracle import pyg_restated as O

F64 = torch.float64


def _random_multigraph(n, e, seed, self_loops=True, isolated=2):
    g = torch.Generator().manual_seed(seed)
    live = n - isolated                                   # the last `isolated` nodes get no edge at all
    src = torch.randint(0, live, (e,), generator=g)
    dst = torch.randint(0, live, (e,), generator=g)
    if not self_loops:
        dst = torch.where(dst == src, (dst + 1) % live, dst)
    ei = torch.stack([src, dst])
    dup = ei[:, : e // 4]                                 # parallel edges
    return torch.cat([ei, dup], dim=1)


def _adjacency(ei, n):
    A = torch.zeros(n, n, dtype=F64)
    A.index_put_((ei[1], ei[0]), torch.ones(ei.size(1), dtype=F64), accumulate=True)   # A[dst, src] = multiplicity
    return A


GRAPHS = [
    pytest.param(dict(n=5, ei=torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]])), id="reference-path-graph"),
    pytest.param(dict(n=40, ei=_random_multigraph(40, 160, 1)), id="multigraph-selfloops-isolated"),
    pytest.param(dict(n=64, ei=_random_multigraph(64, 90, 2, self_loops=False, isolated=9)), id="sparse-no-selfloops"),
    pytest.param(dict(n=7, ei=torch.zeros(2, 0, dtype=torch.long)), id="no-edges"),
    pytest.param(dict(n=6, ei=torch.tensor([[0, 0, 0, 3, 3], [0, 0, 1, 3, 2]])), id="duplicate-selfloops"),
]


def _compare(out_o, out_d, params, x, tol=1e-12):
    torch.testing.assert_close(out_o, out_d, rtol=tol, atol=tol)
    g = torch.Generator().manual_seed(99)
    proj = torch.randn(out_o.shape, generator=g, dtype=F64)
    leaves = [x] + list(params)
    go = torch.autograd.grad((out_o * proj).sum(), leaves, allow_unused=True)
    gd = torch.autograd.grad((out_d * proj).sum(), leaves, allow_unused=True)
    for a, b in zip(go, gd):
        assert (a is None) == (b is None)
        if a is not None:
            torch.testing.assert_close(a, b, rtol=tol, atol=tol)


@pytest.mark.parametrize("graph", GRAPHS)
def test_gcn_conv_is_the_renormalised_adjacency_product(graph):
    n, ei = graph["n"], graph["ei"]
    torch.manual_seed(3)
    conv = O.GCNConv(9, 5).to(F64)
    with torch.no_grad():
        conv.bias.uniform_(-1, 1)
    x = torch.randn(n, 9, dtype=F64, requires_grad=True)

    A = _adjacency(ei, n)
    A = A - torch.diag(torch.diagonal(A)) + torch.eye(n, dtype=F64)       # one self loop of weight 1 per node
    d = A.sum(dim=1)                                                        # in-degree incl. the self loop (>= 1)
    A_hat = A / torch.sqrt(d).view(-1, 1) / torch.sqrt(d).view(1, -1)
    dense = A_hat @ (x @ conv.lin.weight.t()) + conv.bias
    _compare(conv(x, ei), dense, list(conv.parameters()), x)


@pytest.mark.parametrize("graph", GRAPHS)
def test_sage_conv_is_the_row_normalised_adjacency_product(graph):
    n, ei = graph["n"], graph["ei"]
    torch.manual_seed(4)
    conv = O.SAGEConv(9, 5).to(F64)
    x = torch.randn(n, 9, dtype=F64, requires_grad=True)

    A = _adjacency(ei, n)                                                   # self loops and duplicates count as edges
    deg = A.sum(dim=1).clamp(min=1.0)
    dense = ((A @ x) / deg.view(-1, 1)) @ conv.lin_l.weight.t() + conv.lin_l.bias + x @ conv.lin_r.weight.t()
    _compare(conv(x, ei), dense, list(conv.parameters()), x)


@pytest.mark.parametrize("graph", GRAPHS)
@pytest.mark.parametrize("heads,concat", [(4, True), (1, False), (3, False)])
def test_gat_conv_is_the_masked_dense_softmax(graph, heads, concat):
    n, ei = graph["n"], graph["ei"]
    C = 6
    torch.manual_seed(5)
    conv = O.GATConv(9, C, heads=heads, concat=concat).to(F64)
    with torch.no_grad():
        conv.bias.uniform_(-1, 1)
    x = torch.randn(n, 9, dtype=F64, requires_grad=True)

    M = _adjacency(ei, n)
    M = M - torch.diag(torch.diagonal(M)) + torch.eye(n, dtype=F64)       # multiplicities; exactly one self loop
    xs = (x @ conv.lin.weight.t()).view(n, heads, C)
    a_src = (xs * conv.att_src).sum(-1)                                     # [n, H], indexed by the source j
    a_dst = (xs * conv.att_dst).sum(-1)                                     # [n, H], indexed by the destination i
    e = torch.nn.functional.leaky_relu(a_dst.view(n, 1, heads) + a_src.view(1, n, heads), 0.2)   # [i, j, h]
    w = M.view(n, n, 1) * torch.exp(e - e.detach().amax(dim=1, keepdim=True))                     # masked, stabilised
    alpha = w / w.sum(dim=1, keepdim=True)
    out = torch.einsum("ijh,jhc->ihc", alpha, xs)
    dense = (out.reshape(n, heads * C) if concat else out.mean(dim=1)) + conv.bias
    # the oracle stabilises with the max over the row's OWN entries, this statement with the max over all j: both
    # cancel exactly in the quotient up to rounding, and the +1e-16 in the denominator is below 1e-15 relative
    _compare(conv(x, ei), dense, list(conv.parameters()), x, tol=1e-11)


def test_gcn_norm_weights_are_the_entries_of_the_renormalised_adjacency():
    """`gcn_norm` per-edge weights, summed over parallel edges, equal the dense matrix entry for entry."""
    n, ei = 40, _random_multigraph(40, 160, 7)
    ei2, w = O.gcn_norm(ei, n, dtype=F64)
    got = torch.zeros(n, n, dtype=F64)
    got.index_put_((ei2[1], ei2[0]), w, accumulate=True)
    A = _adjacency(ei, n)
    A = A - torch.diag(torch.diagonal(A)) + torch.eye(n, dtype=F64)
    d = A.sum(dim=1)
    torch.testing.assert_close(got, A / torch.sqrt(d).view(-1, 1) / torch.sqrt(d).view(1, -1), rtol=1e-13, atol=1e-13)
    assert ei2.size(1) == int((ei[0] != ei[1]).sum()) + n                   # off-diagonal edges kept + one loop per node
