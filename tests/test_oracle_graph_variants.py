"""CPU tests pinning the graph-variant oracle (`oracle/graph_variants.py`) to the REFERENCE's own functions:
`tests/golden/variants_golden.pt` holds what `src.analysis.hub_ablation.build_edge_index_ablated` and
`src.analysis.robustness.drop_edges` (unmodified, imported from /root/reference by
`tests/golden/make_variants_golden.py`) return on seeded graphs.  The GPU tests (`test_gpu_graph_variants.py`) hold the
device kernels bit-exact to a restatement quoted in that file; here that restatement and the oracle module are shown to
be the same function and to reproduce the reference's outputs, which closes the chain reference -> oracle -> kernel."""
import os

import pytest
import torch

from oracle import graph_variants as V

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "variants_golden.pt"))


def _graph(name):
    g = GOLD["graphs"][name]
    return g["num_nodes"], g["edge_index"].long()


@pytest.mark.parametrize("i", range(len(GOLD["ablate"])))
def test_hub_ablation_oracle_reproduces_the_reference(i):
    rec = GOLD["ablate"][i]
    n, ei = _graph(rec["graph"])
    want = rec["edge_index"].long()
    got, k, hubs, untied = V.ablate_hubs(ei, n, rec["frac"], stable=True)
    assert k == rec["num_hubs"] == int(hubs.sum())
    # the removed-edge COUNT depends on the tie rule only through which tied node is taken; the reference's own
    # torch.topk choice is reproduced exactly whenever the k-th and (k+1)-th degrees differ ...
    if untied:
        assert torch.equal(got, want)
    # ... and by the `stable=False` branch (the reference's very call) always
    got_topk, k2, _, _ = V.ablate_hubs(ei, n, rec["frac"], stable=False)
    assert k2 == rec["num_hubs"] and torch.equal(got_topk, want)
    # survivors keep the original order and never touch a hub
    assert not (hubs[got[0]] | hubs[got[1]]).any()
    if rec["frac"] == 0.0:
        assert torch.equal(got, ei)
    if rec["frac"] == 1.0:
        assert got.size(1) == 0


def test_tied_cases_are_covered_and_resolved_towards_the_lower_id():
    """At least one golden case has a tie at the cut, and there the stable rule takes the lowest ids of the tied
    degree (the rule the CUDA path documents)."""
    tied = 0
    for rec in GOLD["ablate"]:
        n, ei = _graph(rec["graph"])
        _, k, hubs, untied = V.ablate_hubs(ei, n, rec["frac"], stable=True)
        if untied or k == 0:
            continue
        tied += 1
        deg = torch.bincount(ei[0], minlength=n) + torch.bincount(ei[1], minlength=n)
        cut = torch.sort(deg, descending=True).values[k - 1]
        assert hubs[deg > cut].all() and not hubs[deg < cut].any()
        at_cut = torch.nonzero(deg == cut).view(-1)
        taken = int(hubs[at_cut].sum())
        assert hubs[at_cut[:taken]].all() and not hubs[at_cut[taken:]].any()
    assert tied > 0


@pytest.mark.parametrize("i", range(len(GOLD["drop"])))
def test_edge_drop_oracle_reproduces_the_reference(i):
    rec = GOLD["drop"][i]
    n, ei = _graph(rec["graph"])
    torch.manual_seed(rec["seed"])
    if "error" in rec:
        with pytest.raises(RuntimeError, match="empty graph"):
            V.drop_edges(ei, rec["frac"])
        return
    got, cnt = V.drop_edges(ei, rec["frac"])
    assert cnt == rec["count"] == V.drop_count(ei.size(1), rec["frac"])
    assert torch.equal(got, rec["edge_index"].long())
    # the same result with the permutation passed explicitly (how the GPU test drives `egnn_b200.drop_edges`)
    torch.manual_seed(rec["seed"])
    if cnt:
        perm = torch.randperm(ei.size(1))
        got2, _ = V.drop_edges(ei, rec["frac"], perm=perm)
        assert torch.equal(got2, got)


def test_edge_drop_argument_errors():
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]])
    for bad in (-0.1, 1.5):
        with pytest.raises(ValueError):
            V.drop_edges(ei, bad)
    with pytest.raises(RuntimeError):
        V.drop_edges(ei, 1.0)
    assert V.drop_count(8, 0.0625) == 0 and V.drop_count(8, 0.1875) == 2     # Python round: half to even (0.5 -> 0, 1.5 -> 2)


def test_gpu_test_restatement_is_the_oracle():
    """`tests/test_gpu_graph_variants.py:ref_ablate` (what the kernels are held to, bit for bit) == the pinned oracle."""
    import test_gpu_graph_variants as T
    for rec in GOLD["ablate"]:
        n, ei = _graph(rec["graph"])
        for stable in (True, False):
            a = T.ref_ablate(ei, n, rec["frac"], stable)
            b = V.ablate_hubs(ei, n, rec["frac"], stable)
            assert torch.equal(a[0], b[0]) and a[1] == b[1] and torch.equal(a[2], b[2]) and a[3] == b[3]
