"""K1 parity: the sm_100a graph build against the NumPy oracle -- bit-exact integer outputs and
bit-exact gcn_norm weights, on tiny / adversarial / full-size inputs."""
import numpy as np
import pytest
import torch

from oracle import graph_build_np as G

pytestmark = pytest.mark.gpu


def _check(egnn, ei_cpu, n, symmetrize, self_loops):
    g = egnn.build_graph(ei_cpu.cuda(), n, symmetrize=symmetrize, self_loops=self_loops, want_norm=True,
                         keep_edge_list=True, want_order=True)
    ei = ei_cpu.numpy()
    ref = G.symmetrize(ei) if symmetrize else ei
    if self_loops:
        ref2, w, dis = G.gcn_norm(ref, n)
    else:
        ref2 = ref
        deg = np.bincount(ref2[1], minlength=n).astype(np.float32)
        with np.errstate(divide="ignore"):
            dis = (np.float32(1) / np.sqrt(deg)).astype(np.float32)
        dis[np.isinf(dis)] = 0
        w = (dis[ref2[0]] * dis[ref2[1]]).astype(np.float32)
    E2 = g.n_edges
    assert E2 == ref2.shape[1]
    assert int(g.info[1]) == 0
    assert np.array_equal(g.ei2[:, :E2].cpu().numpy(), ref2)
    for by, (p, o, e) in ((1, (g.csr_ptr, g.csr_src, g.csr_eid)), (0, (g.csc_ptr, g.csc_dst, None))):
        rp, ro, re = G.sorted_view(ref2, n, by)
        assert np.array_equal(p.cpu().numpy(), rp)
        assert np.array_equal(o[:E2].cpu().numpy(), ro)
        if e is not None:
            assert np.array_equal(e[:E2].cpu().numpy(), re)
    # csc_pos: CSR position of every CSC entry
    _, _, csr_eid = G.sorted_view(ref2, n, 1)
    _, _, csc_eid = G.sorted_view(ref2, n, 0)
    inv = np.empty(E2, dtype=np.int64)
    inv[csr_eid] = np.arange(E2)
    assert np.array_equal(g.csc_pos[:E2].cpu().numpy(), inv[csc_eid])
    # float outputs: compare BIT patterns
    bits = lambda a: np.asarray(a, dtype=np.float32).view(np.uint32)
    assert np.array_equal(bits(g.dis.cpu().numpy()), bits(dis))
    assert np.array_equal(bits(g.w_edge[:E2].cpu().numpy()), bits(w))
    assert np.array_equal(bits(g.w_csr[:E2].cpu().numpy()), bits(w[csr_eid]))
    assert np.array_equal(bits(g.w_csc[:E2].cpu().numpy()), bits(w[csc_eid]))
    # longest-rows-first schedule inside 32768-row tiles (a stable permutation)
    for ptr_, order in ((g.csr_ptr, g.csr_order), (g.csc_ptr, g.csc_order)):
        d = np.minimum(np.diff(ptr_.cpu().numpy()), 64)
        key = (np.arange(n) >> 15) * 128 + (64 - d)   # descending degree inside 32768-row tiles
        want = np.argsort(key, kind="stable")
        assert np.array_equal(order.cpu().numpy(), want)
    # long-row lists
    for view, ptr_ in ((0, g.csr_ptr), (1, g.csc_ptr)):
        d = np.diff(ptr_.cpu().numpy())
        want = np.nonzero(d > 64)[0]
        k = int(g.info[2 + view])
        got = np.sort((g.csr_long if view == 0 else g.csc_long)[:k].cpu().numpy())
        assert np.array_equal(got, want)
    return g


@pytest.mark.parametrize("symmetrize", [False, True])
@pytest.mark.parametrize("self_loops", [False, True])
def test_path_graph_known_answers(egnn, symmetrize, self_loops):
    ei = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]])
    g = _check(egnn, ei, 5, symmetrize, self_loops)
    if symmetrize and not self_loops:
        assert g.ei2[:, :8].cpu().tolist() == [[0, 1, 2, 3, 1, 2, 3, 4], [1, 2, 3, 4, 0, 1, 2, 3]]
    if self_loops and not symmetrize:
        w = g.w_edge[:9].cpu().numpy().view(np.uint32).tolist()
        assert w == [0x3F3504F3, 0x3EFFFFFF, 0x3EFFFFFF, 0x3EFFFFFF, 0x3F800000] + [0x3EFFFFFF] * 4
    if self_loops and symmetrize:
        w = g.w_edge[:13].cpu().numpy().view(np.uint32).tolist()
        assert w[:8] == [0x3ED105EB, 0x3EAAAAAA, 0x3EAAAAAA, 0x3ED105EB] * 2
        assert w[8:] == [0x3EFFFFFF, 0x3EAAAAAA, 0x3EAAAAAA, 0x3EAAAAAA, 0x3EFFFFFF]


@pytest.mark.parametrize("symmetrize", [False, True])
@pytest.mark.parametrize("self_loops", [False, True])
def test_adversarial(egnn, symmetrize, self_loops):
    from egnn_b200 import synthetic
    gr = synthetic.adversarial_tiny()
    _check(egnn, gr.edge_index, gr.num_nodes, symmetrize, self_loops)


def test_empty_and_single(egnn):
    _check(egnn, torch.zeros((2, 0), dtype=torch.int64), 7, False, True)
    _check(egnn, torch.zeros((2, 0), dtype=torch.int64), 7, True, False)
    _check(egnn, torch.tensor([[0], [0]]), 1, True, True)


def test_random_multigraph(egnn):
    g = torch.Generator().manual_seed(1)
    n = 3001
    ei = torch.randint(0, n, (2, 50_000), generator=g)  # duplicates and self loops galore
    for s in (False, True):
        for l in (False, True):
            _check(egnn, ei, n, s, l)


def test_out_of_range_raises(egnn):
    ei = torch.tensor([[0, 9], [1, 2]]).cuda()
    with pytest.raises(IndexError):
        egnn.build_graph(ei, 5)


def test_cpu_tensor_is_an_error(egnn):
    with pytest.raises(RuntimeError):
        egnn.build_graph(torch.tensor([[0], [1]]), 2)


def test_full_size_elliptic(egnn):
    from egnn_b200 import synthetic
    gr = synthetic.make_elliptic_like()
    _check(egnn, gr.edge_index, gr.num_nodes, True, False)
    _check(egnn, gr.edge_index, gr.num_nodes, False, True)
    sym = egnn.symmetrize(gr.edge_index.cuda())
    assert torch.equal(sym.cpu(), torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1))


def test_cache_is_keyed(egnn):
    from egnn_b200.graph import GraphCache
    c = GraphCache()
    a = torch.tensor([[0, 1], [1, 2]]).cuda()
    b = torch.tensor([[0, 1], [1, 0]]).cuda()
    ga, gb = c.get(a, 3), c.get(b, 3)
    assert ga is not gb and c.get(a, 3) is ga and c.get(a, 3, self_loops=True) is not ga
    a[0, 0] = 2  # in-place edit bumps the version -> rebuilt
    assert c.get(a, 3) is not ga
