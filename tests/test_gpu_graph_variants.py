"""SURVEY 8(f) rank 2 -- re-entrant graph variants built on the device: hub ablation (`src/train_gnn.py:526-540`,
`src/analysis/hub_ablation.py:56-71`) and random edge drop (`src/analysis/robustness.py:65-82`), bit-exact against the
reference's torch code restated below (its modules import torch_geometric, so the 10 lines are quoted, not imported)."""
import pytest
import torch

from oracle import pyg_restated as O
from util import REL_FP32, assert_close

pytestmark = pytest.mark.gpu


def ref_ablate(edge_index, num_nodes, frac, stable):
    """hub_ablation.build_edge_index_ablated; `stable`: resolve topk ties towards the lower node id"""
    num_hubs = int(float(frac) * float(num_nodes))
    ei = edge_index.detach().cpu()
    deg = torch.bincount(ei[0], minlength=num_nodes) + torch.bincount(ei[1], minlength=num_nodes)
    hubs = torch.zeros(num_nodes, dtype=torch.bool)
    untied = True
    if num_hubs > 0:
        idx = (torch.sort(deg, descending=True, stable=True).indices[:num_hubs] if stable
               else torch.topk(deg, num_hubs).indices)
        hubs[idx] = True
        srt = torch.sort(deg, descending=True).values
        untied = num_hubs >= num_nodes or bool(srt[num_hubs - 1] != srt[num_hubs])
    mask = ~(hubs[ei[0]] | hubs[ei[1]])
    return ei[:, mask], num_hubs, hubs, untied


def _graphs():
    from egnn_b200 import synthetic
    adv = synthetic.adversarial_tiny()
    small = synthetic.make_elliptic_like(n_nodes=6000, n_edges=7000, n_timesteps=12, seed=3, hub_degree=200)
    full = synthetic.make_elliptic_like()
    sym = lambda g: torch.cat([g.edge_index, g.edge_index.flip(0)], 1)
    return {"adversarial": (adv.num_nodes, adv.edge_index), "small_sym": (small.num_nodes, sym(small)),
            "full": (full.num_nodes, full.edge_index), "full_sym": (full.num_nodes, sym(full))}


@pytest.mark.parametrize("gname", ["adversarial", "small_sym", "full", "full_sym"])
@pytest.mark.parametrize("frac", [0.0, 0.0005, 0.01, 0.05, 0.5])
def test_hub_ablation_bit_exact(egnn, gname, frac):
    n, ei = _graphs()[gname]
    got, num_hubs, hub = egnn.ablate_hubs(ei.cuda(), n, frac)
    want, want_hubs, hubs_ref, untied = ref_ablate(ei, n, frac, stable=True)
    assert num_hubs == want_hubs and int(hub.sum()) == want_hubs
    assert torch.equal(hub.cpu(), hubs_ref)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)          # same edges, same (original) order
    if untied:   # the k-th and (k+1)-th largest degrees differ: the reference's own torch.topk picks the same set
        want_topk, _, hubs_topk, _ = ref_ablate(ei, n, frac, stable=False)
        assert torch.equal(hub.cpu(), hubs_topk) and torch.equal(got.cpu(), want_topk)


def test_hub_ablation_edge_cases(egnn):
    none = torch.zeros((2, 0), dtype=torch.int64, device="cuda")
    got, k, hub = egnn.ablate_hubs(none, 7, 0.5)
    assert got.shape == (2, 0) and k == 3 and int(hub.sum()) == 3 and hub[:3].all()    # all degrees tie at 0: ids 0,1,2
    with pytest.raises(IndexError):
        egnn.ablate_hubs(torch.tensor([[0, 9], [1, 2]], device="cuda"), 5, 0.2)
    with pytest.raises(RuntimeError):
        egnn.ablate_hubs(torch.tensor([[0], [1]]), 5, 0.2)                             # CPU tensor: no fallback


@pytest.mark.parametrize("drop_frac", [0.0, 0.1, 0.5, 0.999])
def test_edge_drop_bit_exact_for_the_same_permutation(egnn, drop_frac):
    from egnn_b200 import synthetic
    gr = synthetic.make_elliptic_like()
    ei = gr.edge_index
    E = ei.size(1)
    perm = torch.randperm(E, generator=torch.Generator().manual_seed(5))
    got, cnt = egnn.drop_edges(ei.cuda(), drop_frac, perm=perm.cuda())
    drop_count = min(int(round(drop_frac * float(E))), E)            # robustness.drop_edges, restated
    want = ei if drop_count == 0 else ei[:, perm[drop_count:]]
    assert cnt == drop_count and torch.equal(got.cpu(), want)
    with pytest.raises(ValueError):
        egnn.drop_edges(ei.cuda(), 1.5)
    with pytest.raises(RuntimeError):
        egnn.drop_edges(ei.cuda(), 1.0)
    auto, cnt2 = egnn.drop_edges(ei.cuda(), 0.25)                    # own randperm: right count, a subset of the edges
    assert cnt2 == int(round(0.25 * E)) and auto.size(1) == E - cnt2


def test_model_on_ablated_graph_matches_oracle(egnn, small_graph):
    """The re-entrant call of the reference (`get_probs(edge_index_abl)`, src/train_gnn.py:541): the same model, a new
    edge tensor -> its own cached structure, logits equal to the oracle's on the reference-ablated edge list."""
    gr = small_graph
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    cfg = dict(arch="sage_resbn", hidden_dim=64, layers=3, dropout=0.2, time_embed_dim=2, time_embed_type="sin",
               max_timestep=49)
    torch.manual_seed(0)
    ours = egnn.build_model("sage_resbn", 166, cfg)
    ref = O.build_model("sage_resbn", 166, cfg)
    ref.load_state_dict(ours.state_dict())
    ours = ours.cuda().eval()
    ref.eval()
    abl, k, _ = egnn.ablate_hubs(ei.cuda(), gr.num_nodes, 0.02)
    want_ei, _, _, _ = ref_ablate(ei, gr.num_nodes, 0.02, stable=True)
    assert torch.equal(abl.cpu(), want_ei) and k == int(0.02 * gr.num_nodes)
    with torch.no_grad():
        for e_ours, e_ref in ((ei.cuda(), ei), (abl, want_ei)):
            lo = ours(gr.x.cuda(), e_ours, gr.timestep.cuda())
            lr_ = ref(gr.x, e_ref, gr.timestep)
            assert_close(lo, lr_, REL_FP32, "logits")
