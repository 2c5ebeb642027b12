"""K2/K3 parity: SpMM forward / transposed backward against the CPU scatter oracle.
fp32 results must be BITWISE equal (sequential edge-order accumulation)."""
import numpy as np
import pytest
import torch

from oracle import pyg_restated as O

pytestmark = pytest.mark.gpu


def _graphs():
    from egnn_b200 import synthetic
    return {
        "adv": synthetic.adversarial_tiny(),
        "small": synthetic.make_elliptic_like(n_nodes=6000, n_edges=7000, n_timesteps=12, seed=3, hub_degree=200),
    }


def _feat(n, f, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, f, generator=g)
    x[:, ::7] *= 50.0  # wide dynamic range so summation order shows up in the bits
    return x


@pytest.mark.parametrize("gname", ["adv", "small"])
@pytest.mark.parametrize("F", [2, 4, 12, 32, 64, 128, 167, 168, 300])
def test_mean_fwd_and_bwd_bitexact_fp32(egnn, gname, F):
    from egnn_b200 import ops, _lib
    gr = _graphs()[gname]
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1)
    n = gr.num_nodes
    x = _feat(n, F).requires_grad_(True)
    ref = O.scatter_mean(x.index_select(0, ei[0]), ei[1], n)
    gout = _feat(n, F, seed=5)
    ref.backward(gout)
    g = egnn.build_graph(ei.cuda(), n)
    got = ops.spmm(g, "csr", _lib.SPMM_MEAN, x.detach().cuda(), torch.float32)
    assert torch.equal(got.cpu(), ref.detach()), "forward mean not bit-exact"
    gx = ops.spmm(g, "csc", _lib.SPMM_DIV_NBR, gout.cuda(), torch.float32)
    assert torch.equal(gx.cpu(), x.grad), "transposed backward not bit-exact"
    # accumulate epilogue
    base = _feat(n, F, seed=9).cuda()
    acc = base.clone()
    ops.spmm(g, "csc", _lib.SPMM_DIV_NBR, gout.cuda(), torch.float32, out=acc, accumulate=True)
    assert torch.equal(acc.cpu(), base.cpu() + x.grad)


@pytest.mark.parametrize("F", [2, 32, 128, 130])
def test_gcn_weighted_bitexact_fp32(egnn, F):
    from egnn_b200 import ops, _lib
    gr = _graphs()["small"]
    n = gr.num_nodes
    ei = gr.edge_index
    h = _feat(n, F).requires_grad_(True)
    bias = torch.randn(F)
    ei2, w = O.gcn_norm(ei, n)
    ref = O.scatter_sum(w.view(-1, 1) * h.index_select(0, ei2[0]), ei2[1], n) + bias
    gout = _feat(n, F, seed=2)
    ref.backward(gout)
    g = egnn.build_graph(ei.cuda(), n, self_loops=True)
    got = ops.spmm(g, "csr", _lib.SPMM_WEIGHTED, h.detach().cuda(), torch.float32, bias=bias.cuda())
    assert torch.equal(got.cpu(), ref.detach())
    gh = ops.spmm(g, "csc", _lib.SPMM_WEIGHTED, gout.cuda(), torch.float32)
    assert torch.equal(gh.cpu(), h.grad)


@pytest.mark.parametrize("F", [64, 168])
def test_bf16_paths(egnn, F):
    """bf16 in/out with fp32 accumulation: compare against the fp32 oracle of the bf16-rounded input;
    the only error left is the final rounding (half an ulp of bf16 = 2^-9 relative)."""
    from egnn_b200 import ops, _lib
    gr = _graphs()["small"]
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1)
    n = gr.num_nodes
    xb = _feat(n, F).bfloat16()
    ref = O.scatter_mean(xb.float().index_select(0, ei[0]), ei[1], n)
    g = egnn.build_graph(ei.cuda(), n)
    got = ops.spmm(g, "csr", _lib.SPMM_MEAN, xb.cuda(), torch.bfloat16)
    # bf16 output: the mean is sum * rn(1/deg) (not an IEEE division) before the final rounding, so
    # a result can land on the neighbouring bf16 value when the quotient sits on a rounding boundary
    assert (got.cpu().float() - ref).abs().max() <= ref.abs().max() * 2 ** -8
    assert (got.cpu() != ref.bfloat16()).float().mean() < 1e-3
    got32 = ops.spmm(g, "csr", _lib.SPMM_MEAN, xb.cuda(), torch.float32)
    assert torch.equal(got32.cpu(), ref)
    gotfb = ops.spmm(g, "csr", _lib.SPMM_MEAN, xb.float().cuda(), torch.bfloat16)
    assert (gotfb.cpu() != ref.bfloat16()).float().mean() < 1e-3


def test_determinism_and_full_size(egnn):
    from egnn_b200 import ops, _lib, synthetic
    gr = synthetic.make_elliptic_like()
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1).cuda()
    g = egnn.build_graph(ei, gr.num_nodes)
    x = torch.nn.functional.pad(gr.x, (0, 2)).cuda()
    a = ops.spmm(g, "csr", _lib.SPMM_MEAN, x, torch.float32)
    b = ops.spmm(g, "csr", _lib.SPMM_MEAN, x, torch.float32)
    assert torch.equal(a, b)
    # size-independent properties: linearity and the constant-vector fixed point
    ones = torch.ones_like(x)
    m1 = ops.spmm(g, "csr", _lib.SPMM_MEAN, ones, torch.float32)
    deg = (g.csr_ptr[1:] - g.csr_ptr[:-1]).float()
    assert torch.equal(m1[:, 0], (deg > 0).float())
    s = ops.spmm(g, "csr", _lib.SPMM_SUM, ones, torch.float32)
    assert torch.equal(s[:, 0], deg)
    # transposed sum of ones = out-degree; total mass conserved
    st = ops.spmm(g, "csc", _lib.SPMM_SUM, ones, torch.float32)
    assert float(st[:, 0].sum()) == float(s[:, 0].sum()) == float(g.n_edges)
    # every row against the CPU oracle (lane groups of the streaming kernel span several partition tasks here)
    ref = O.scatter_mean(x.cpu().index_select(0, ei[0].cpu()), ei[1].cpu(), gr.num_nodes)
    assert torch.equal(a.cpu(), ref)


def test_full_size_bf16_and_weighted_bitexact(egnn):
    """Full Elliptic-shaped graph: the narrow bf16 configuration (4 lanes x 2 vectors per row) and the GCN-weighted
    mode of the streaming kernel, fp32 accumulation compared bit for bit with the oracle."""
    from egnn_b200 import ops, _lib, synthetic
    gr = synthetic.make_elliptic_like()
    n = gr.num_nodes
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1)
    g = egnn.build_graph(ei.cuda(), n)
    xb = _feat(n, 64).bfloat16()
    ref = O.scatter_mean(xb.float().index_select(0, ei[0]), ei[1], n)
    got = ops.spmm(g, "csr", _lib.SPMM_MEAN, xb.cuda(), torch.float32)
    assert torch.equal(got.cpu(), ref)
    gotb = ops.spmm(g, "csr", _lib.SPMM_MEAN, xb.cuda(), torch.bfloat16)
    assert (gotb.cpu().float() - ref).abs().max() <= 2.0 ** -8 * ref.abs().max()   # one bf16 rounding
    # transposed sum + addend (the backward's accumulate into the root-path gradient)
    base = _feat(n, 64, seed=3)
    reft = O.scatter_sum(xb.float().index_select(0, ei[1]), ei[0], n)
    gott = ops.spmm(g, "csc", _lib.SPMM_SUM, xb.cuda(), torch.float32, addend=base.cuda())
    assert torch.equal(gott.cpu(), base + reft)
    # GCN: weights D^-1/2 (A+I) D^-1/2 on the self-loop graph, bias epilogue
    h = _feat(n, 128, seed=4)
    bias = torch.randn(128)
    ei2, w = O.gcn_norm(gr.edge_index, n)
    refw = O.scatter_sum(w.view(-1, 1) * h.index_select(0, ei2[0]), ei2[1], n) + bias
    gl = egnn.build_graph(gr.edge_index.cuda(), n, self_loops=True)
    gotw = ops.spmm(gl, "csr", _lib.SPMM_WEIGHTED, h.cuda(), torch.float32, bias=bias.cuda())
    assert torch.equal(gotw.cpu(), refw)


def test_row_partition_matches_definition(egnn):
    """egnn_spmm_partition: part[k] = first row r with r + 2 * ptr[r] >= 32 k (n_rows past the end)."""
    import numpy as np
    from egnn_b200 import synthetic
    for gr in _graphs().values():
        ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1)
        g = egnn.build_graph(ei.cuda(), gr.num_nodes)
        for ptr_, part in ((g.csr_ptr, g.csr_part), (g.csc_ptr, g.csc_part)):
            p = ptr_.cpu().numpy().astype(np.int64)
            cost = np.arange(p.size, dtype=np.int64) + 2 * p          # non-decreasing in r
            want = np.searchsorted(cost, 32 * np.arange(g.n_tasks + 1, dtype=np.int64), side="left")
            want = np.minimum(want, gr.num_nodes)
            got = part.cpu().numpy()
            assert got.shape[0] == g.n_tasks + 1 and got[0] == 0 and got[-1] == gr.num_nodes
            assert np.array_equal(got, want)


def test_large_graph_groups_of_eight_tasks_bitexact(egnn):
    """8x replicated graph (1.6 M rows): the streaming kernel hands out groups of 8 partition tasks here (more tasks
    than two waves of lane groups), unlike on the base graph.  fp32 mean aggregation bit for bit against the oracle,
    plus the size-independent checks."""
    from egnn_b200 import ops, _lib, synthetic
    gr = synthetic.replicate(synthetic.make_elliptic_like(), 8)
    n = gr.num_nodes
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1)
    g = egnn.build_graph(ei.cuda(), n)
    assert g.n_tasks // 8 > 4 * 148 * 16 * 2          # the large-matrix branch of the launcher (4 CTAs x 16 groups)
    x = _feat(n, 64, seed=7)
    got = ops.spmm(g, "csr", _lib.SPMM_MEAN, x.cuda(), torch.float32).cpu()
    ref = O.scatter_mean(x.index_select(0, ei[0]), ei[1], n)
    assert torch.equal(got, ref)
    ones = torch.ones(n, 64, device="cuda")
    s = ops.spmm(g, "csc", _lib.SPMM_SUM, ones, torch.float32)
    deg_out = (g.csc_ptr[1:] - g.csc_ptr[:-1]).float()
    assert torch.equal(s[:, 0], deg_out) and float(s[:, 0].sum()) == float(g.n_edges)
