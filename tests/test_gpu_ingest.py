"""SURVEY 8(f) rank 3 -- ingestion on the device: `egnn_txid_join` / `egnn_temporal_masks` bit-exact against the
outputs of the reference's own loader (tests/golden/ingest_golden.pt) and, at full size, against the oracle and a
round-trip property (relabel the synthetic graph with random txIds, shuffle in junk edges, join -> the original
edge list comes back)."""
import os

import numpy as np
import pytest
import torch

from oracle import ingest_np as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_golden.pt")


@pytest.fixture(scope="module")
def cases():
    return torch.load(GOLD, weights_only=False)


def test_join_bitexact_vs_reference_loader(cases):
    from egnn_b200 import ingest
    for c in cases:
        ei, cnt = ingest.join_edges(c["tx_ids"].cuda(), c["timestep"].cuda(), c["e_src_tx"].cuda(), c["e_dst_tx"].cuda())
        assert ei.dtype == torch.int64 and ei.is_contiguous()
        assert torch.equal(ei.cpu(), c["edge_index"]), c["spec"]
        assert cnt["kept"] == c["meta"]["num_edges"] and f"mapped={cnt['mapped']} " in c["log"]
        assert cnt["duplicate_txids"] == (1 if c["spec"]["dup_tx"] else 0)


def test_masks_bitexact_vs_reference(cases):
    from egnn_b200 import ingest
    from egnn_b200.synthetic import EllipticGraph
    for c in cases:
        for name, m in c["masks"].items():
            a, b, k = m["args"]
            d = EllipticGraph(x=c["x"].cuda(), edge_index=c["edge_index"].cuda(), y=c["y"].cuda(), timestep=c["timestep"].cuda())
            assert ingest.make_temporal_masks(d, a, b, k) is d
            for key in ("train", "val", "test"):
                got = getattr(d, key + "_mask")
                assert got.dtype == torch.bool and torch.equal(got.cpu(), m[key]), (c["spec"], name, key)


def test_loader_end_to_end_from_csv(cases, tmp_path, capsys):
    """The public entry point on the very CSV text the reference loader read: same x / y / timestep / edge_index /
    meta and the same [EDGES] log line."""
    from egnn_b200 import ingest
    for i, c in enumerate(cases):
        d = tmp_path / f"case{i}"
        os.makedirs(d)
        for fn, text in c["files"].items():
            (d / fn).write_text(text)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            data, meta = ingest.load_elliptic_as_graph(str(d))
        out = capsys.readouterr().out
        assert data.x.is_cuda and torch.equal(data.x.cpu(), c["x"])
        assert torch.equal(data.y.cpu(), c["y"]) and torch.equal(data.timestep.cpu(), c["timestep"])
        assert torch.equal(data.edge_index.cpu(), c["edge_index"])
        assert meta == c["meta"]
        assert out.strip().splitlines()[-1] == c["log"].strip().splitlines()[-1]


def test_full_size_round_trip_and_forward():
    import egnn_b200 as E
    from egnn_b200 import ingest, synthetic
    gr = synthetic.make_elliptic_like()                      # 203 769 nodes, 234 355 edges
    N, Eg = gr.num_nodes, gr.edge_index.size(1)
    g = torch.Generator().manual_seed(5)
    tx = torch.randperm(N * 8, generator=g)[:N] * 1_000_003 + 17     # distinct, sparse 64-bit ids
    tx[7] = -1                                                # the all-ones key takes the side slot
    src_tx, dst_tx = tx[gr.edge_index[0]], tx[gr.edge_index[1]]
    # junk rows: unknown endpoints and cross-timestep pairs, scattered through the list
    n_junk = 20_000
    a = torch.randint(0, N, (n_junk,), generator=g)
    b = torch.randint(0, N, (n_junk,), generator=g)
    junk_s, junk_d = tx[a].clone(), tx[b].clone()
    junk_s[::2] = 3                                           # 3 is not a multiple of 1 000 003 plus 17
    keep_junk = (gr.timestep[a] == gr.timestep[b]) & (torch.arange(n_junk) % 2 == 1)   # same-timestep junk survives
    pos = torch.sort(torch.randint(0, Eg + 1, (n_junk,), generator=g)).values
    es = torch.empty(Eg + n_junk, dtype=torch.int64)
    ed = torch.empty_like(es)
    is_junk = torch.zeros(Eg + n_junk, dtype=torch.bool)
    is_junk[pos + torch.arange(n_junk)] = True
    es[is_junk], ed[is_junk] = junk_s, junk_d
    es[~is_junk], ed[~is_junk] = src_tx, dst_tx
    ei, cnt = ingest.join_edges(tx.cuda(), gr.timestep.cuda(), es.cuda(), ed.cuda())
    ref, mapped, kept = O.join_edges(tx.numpy(), gr.timestep.numpy(), es.numpy(), ed.numpy())
    assert torch.equal(ei.cpu(), torch.from_numpy(ref)) and cnt["mapped"] == mapped and cnt["kept"] == kept
    # property: dropping the surviving junk columns returns the generator's edge list, order intact
    surv = torch.ones(Eg + n_junk, dtype=torch.bool)
    surv[is_junk] = keep_junk
    back = ei.cpu()[:, ~is_junk[surv]]
    assert torch.equal(back, gr.edge_index)
    # and the joined graph drives the model like the generator's own tensors do
    d, ei_dev, graph = ingest.to_device_graph(
        synthetic.EllipticGraph(x=gr.x, edge_index=back, y=gr.y, timestep=gr.timestep), symmetrize_edges=True)
    assert graph.n_nodes == N and ei_dev.size(1) == 2 * Eg
    cfg = dict(hidden_dim=64, layers=3, dropout=0.0, time_embed_dim=2, time_embed_type="sin", max_timestep=49)
    torch.manual_seed(0)
    model = E.build_model("sage_resbn", 166, cfg).cuda().eval()
    with torch.no_grad():
        logits = model(d.x, ei_dev, d.timestep)
    assert logits.shape == (N, 2) and torch.isfinite(logits).all()
