"""Mini-batch path: the device-side NeighborLoader (`csrc/sampler.cu`, `egnn_b200/loader.py`) against
the sequential restatement of PyG's sampler (`oracle/neighbor_sample_np.py`) -- node lists, local edge lists and edge
ids bit-exact -- and the property that pins the whole path to the full-batch one: with full fan-outs over as many hops as
the net has layers, the seeds' logits on the sampled subgraph equal their full-graph logits."""
import numpy as np
import pytest
import torch

from oracle import pyg_restated as O
from oracle.neighbor_sample_np import csc_by_destination, neighbor_sample
from util import REL_FP32, assert_close

pytestmark = pytest.mark.gpu


class _Data:
    def __init__(self, gr, edge_index=None):
        self.x, self.y, self.timestep = gr.x, gr.y, gr.timestep
        self.train_mask, self.val_mask, self.test_mask = gr.train_mask, gr.val_mask, gr.test_mask
        self.edge_index = gr.edge_index if edge_index is None else edge_index


def _check_batch(batch, data, ip, src, eid, seeds, fan, seed, batch_idx):
    n_id, le, e_id, nodes, edges = neighbor_sample(ip, src, eid, seeds, fan, seed=seed, batch_idx=batch_idx)
    assert batch.batch_size == len(seeds)
    assert batch.n_id.cpu().numpy().tolist() == n_id.tolist()
    assert batch.edge_index.dtype == torch.int64 and batch.edge_index.is_contiguous()
    assert np.array_equal(batch.edge_index.cpu().numpy(), le)
    assert np.array_equal(batch.e_id.cpu().numpy(), e_id)
    H = len(fan)
    assert batch.counts.cpu().tolist() == nodes + edges and len(nodes) == H + 1
    for k in ("x", "y", "timestep", "train_mask", "val_mask", "test_mask"):
        assert torch.equal(getattr(batch, k).cpu(), getattr(data, k)[torch.from_numpy(n_id)]), k


@pytest.mark.parametrize("fan", [[3, 2], [10, 10], [-1, -1], [1], [2, 2, 2], [500, 1]])
def test_sampler_bit_exact_small(egnn, small_graph, fan):
    gr = small_graph
    sym = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    for ei in (gr.edge_index, sym):
        data = _Data(gr, ei)
        ip, src, eid = csc_by_destination(ei.numpy(), gr.num_nodes)
        idx = torch.nonzero(gr.train_mask).view(-1)
        loader = egnn.NeighborLoader(data, num_neighbors=fan, batch_size=257, input_nodes=idx, shuffle=False, seed=5)
        assert len(loader) == -(-idx.numel() // 257)
        seen = []
        for b, batch in enumerate(loader):
            seeds = idx[b * 257:(b + 1) * 257].numpy()
            assert torch.equal(batch.input_id.cpu(), torch.arange(b * 257, b * 257 + len(seeds)))
            if b < 3 or b == len(loader) - 1:
                _check_batch(batch.to("cuda"), data, ip, src, eid, seeds, fan, 5, b)
            seen.append(batch.n_id[:batch.batch_size].cpu())
        assert torch.equal(torch.cat(seen), idx)                 # every input node is a seed exactly once


def test_sampler_adversarial_and_edge_cases(egnn):
    from egnn_b200 import synthetic
    adv = synthetic.adversarial_tiny()                            # duplicates, self-loops, isolated nodes, a hub
    data = _Data(adv)
    ip, src, eid = csc_by_destination(adv.edge_index.numpy(), adv.num_nodes)
    for fan in ([2, 2], [-1], [64, 3]):
        loader = egnn.NeighborLoader(data, num_neighbors=fan, batch_size=7, shuffle=False, seed=1)
        for b, batch in enumerate(loader):
            seeds = np.arange(b * 7, min((b + 1) * 7, adv.num_nodes))
            _check_batch(batch, data, ip, src, eid, seeds, fan, 1, b)
    # a graph without edges: every batch is its seeds
    empty = _Data(adv, torch.zeros((2, 0), dtype=torch.int64))
    for batch in egnn.NeighborLoader(empty, num_neighbors=[5, 5], batch_size=4):
        assert batch.edge_index.shape == (2, 0) and batch.num_nodes == batch.batch_size
    with pytest.raises(ValueError):
        egnn.NeighborLoader(data, num_neighbors=[2], batch_size=4, input_nodes=torch.tensor([1, 1, 2]))
    with pytest.raises(ValueError):
        egnn.NeighborLoader(data, num_neighbors=[2], batch_size=4, device="cpu")
    with pytest.raises(NotImplementedError):
        egnn.NeighborLoader(data, num_neighbors=[2], batch_size=4, replace=True)


def test_shuffle_is_a_permutation_and_changes_per_epoch(egnn, small_graph):
    data = _Data(small_graph)
    idx = torch.nonzero(small_graph.val_mask).view(-1)
    loader = egnn.NeighborLoader(data, num_neighbors=[4, 4], batch_size=100, input_nodes=idx, shuffle=True, seed=3)
    epochs = []
    for _ in range(2):
        epochs.append(torch.cat([b.n_id[:b.batch_size].cpu() for b in loader]))
    assert torch.equal(torch.sort(epochs[0]).values, idx) and torch.equal(torch.sort(epochs[1]).values, idx)
    assert not torch.equal(epochs[0], epochs[1])


def test_full_size_batch_bit_exact(egnn):
    """The reference's defaults (fanout [10, 10], batch_size 8192, `src/train_gnn.py:333-334`) on the full graph."""
    from egnn_b200 import synthetic
    gr = synthetic.make_elliptic_like(train_window_k=8)
    sym = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    data = _Data(gr, sym)
    ip, src, eid = csc_by_destination(sym.numpy(), gr.num_nodes)
    idx = torch.nonzero(gr.train_mask).view(-1)
    loader = egnn.NeighborLoader(data, num_neighbors=[10, 10], batch_size=8192, input_nodes=idx, shuffle=False, seed=42)
    it = iter(loader)
    for b in range(2):
        batch = next(it)
        _check_batch(batch, data, ip, src, eid, idx[b * 8192:(b + 1) * 8192].numpy(), [10, 10], 42, b)
        deg_in = torch.bincount(batch.edge_index[1], minlength=batch.num_nodes)
        assert int(deg_in.max()) <= 10


@pytest.mark.parametrize("dtype,width", [(torch.float32, 166), (torch.float32, 168), (torch.float32, 1), (torch.int64, 1),
                                         (torch.bool, 1), (torch.bfloat16, 64), (torch.uint8, 3)])
def test_gather_rows(egnn, dtype, width):
    from egnn_b200.loader import gather_rows
    g = torch.Generator().manual_seed(0)
    t = (torch.randn(1000, width, generator=g) * 50).to(dtype)
    if width == 1:
        t = t.view(-1)
    idx = torch.randint(0, 1000, (777,), generator=g)
    assert torch.equal(gather_rows(t.cuda(), idx.cuda()).cpu(), t[idx])
    assert gather_rows(t.cuda(), idx[:0].cuda()).shape[0] == 0


def test_full_fanout_minibatch_equals_full_batch(egnn, small_graph):
    """2-layer SAGE, fan-outs [-1, -1]: the sampled subgraph contains the complete 2-hop in-neighbourhood of every seed,
    in-edges in their original order, so the seeds' logits match the full-graph forward (and the CPU oracle run on the
    sampled subgraph) at the fp32 bar; then one epoch of `train_epoch_minibatch` against the oracle's loop, batch by batch."""
    from egnn_b200.train import train_epoch_minibatch
    gr = small_graph
    sym = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    data = _Data(gr, sym)
    cfg = dict(hidden_dim=64, layers=2, dropout=0.0)
    torch.manual_seed(0)
    ours = egnn.build_model("sage", 166, cfg)
    ref = O.build_model("sage", 166, cfg)
    ref.load_state_dict(ours.state_dict())
    ours = ours.cuda()
    idx = torch.nonzero(gr.train_mask).view(-1)
    ours.eval()
    with torch.no_grad():
        full = ours(gr.x.cuda(), sym.cuda(), None)
    loader = egnn.NeighborLoader(data, num_neighbors=[-1, -1], batch_size=300, input_nodes=idx, shuffle=True, seed=8)
    for batch in loader:
        with torch.no_grad():
            lg = ours(batch.x, batch.edge_index, None)
        seeds = batch.n_id[:batch.batch_size]
        assert_close(lg[:batch.batch_size], full[seeds], REL_FP32, "mini-batch vs full-batch logits")
        lr_ = ref.eval()(batch.x.cpu(), batch.edge_index.cpu(), None)
        assert_close(lg, lr_, REL_FP32, "sampled-subgraph logits vs oracle")
    # the training loop: same batches (same loader seed), same Adam, oracle on the CPU
    cw = O.class_weight(gr.y[gr.train_mask])
    loss_fn = egnn.make_loss_fn({}, cw, ours, 1, 12)
    tcfg = {"grad_clip": 1.0}
    opt_o = torch.optim.Adam(ours.parameters(), lr=1e-3, weight_decay=5e-4)
    opt_r = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=5e-4)
    mk = lambda: egnn.NeighborLoader(data, num_neighbors=[5, 5], batch_size=300, input_nodes=idx, shuffle=True, seed=21)
    loss_o = train_epoch_minibatch(ours, mk(), opt_o, loss_fn, tcfg)
    ref.train()
    tot, cnt = 0.0, 0
    for batch in mk():
        opt_r.zero_grad(set_to_none=True)
        lg = ref(batch.x.cpu(), batch.edge_index.cpu(), None)
        bs = batch.batch_size
        y = batch.y[:bs].cpu()
        loss = O.masked_weighted_ce(lg[:bs], y, torch.ones(bs, dtype=torch.bool), cw)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        opt_r.step()
        tot, cnt = tot + float(loss.detach()) * bs, cnt + bs
    assert abs(loss_o - tot / cnt) <= 1e-4 * abs(tot / cnt), (loss_o, tot / cnt)
    for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        assert_close(p.detach().cpu(), q.detach(), 1e-3, f"parameters after one mini-batch epoch: {n}")


def test_eval_val_minibatch_matches_full_graph_eval(egnn, small_graph):
    """`eval_val_minibatch` (`src/train_gnn.py:258-277`) with full fan-outs = the `eval_split` probabilities of the same
    nodes (SAGE-ResBN in eval mode: BatchNorm on its running statistics, so a sampled subgraph changes nothing)."""
    from egnn_b200.train import eval_probs, eval_val_minibatch
    gr = small_graph
    sym = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    data = _Data(gr, sym)
    cfg = dict(hidden_dim=64, layers=3, dropout=0.2, time_embed_dim=2, time_embed_type="sin", max_timestep=12)
    torch.manual_seed(0)
    model = egnn.build_model("sage_resbn", 166, cfg).cuda()
    idx = torch.nonzero(gr.val_mask).view(-1)
    loader = egnn.NeighborLoader(data, num_neighbors=[-1, -1, -1], batch_size=128, input_nodes=idx, shuffle=False)
    y, p = eval_val_minibatch(model, loader)
    probs_full, _ = eval_probs(model, gr.x.cuda(), sym.cuda(), gr.timestep.cuda())
    assert y.shape == p.shape == (idx.numel(),) and (y == gr.y[idx].numpy()).all()
    assert_close(torch.from_numpy(p), probs_full[idx.cuda()].cpu(), REL_FP32, "mini-batch eval probabilities")
    yd, pd = eval_val_minibatch(model, loader, as_numpy=False)
    assert yd.is_cuda and torch.equal(pd.cpu(), torch.from_numpy(p))



def test_sampler_matches_the_committed_golden_vectors(egnn):
    """tests/golden/neighbor_sample_golden.json (made by make_neighbor_sample_golden.py from the oracle): pins the Philox
    keying, the Floyd draw and the relabelling order of the kernel against a committed file, not only against the oracle
    of the day."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "neighbor_sample_golden.json")))
    for name, g in gold.items():
        n = g["num_nodes"]

        class D:
            pass
        d = D()
        d.x = torch.zeros(n, 4)
        d.y = torch.arange(n)
        d.timestep = torch.ones(n, dtype=torch.int64)
        d.edge_index = torch.tensor(g["edge_index"], dtype=torch.int64)
        seeds = torch.tensor(g["seeds"], dtype=torch.int64, device="cuda")
        for c in g["cases"]:
            loader = egnn.NeighborLoader(d, num_neighbors=c["fanouts"], batch_size=len(g["seeds"]), seed=c["seed"])
            b = loader.sample(seeds, c["batch_idx"])
            assert b.n_id.cpu().tolist() == c["n_id"], (name, c["fanouts"])
            assert b.edge_index.cpu().tolist() == c["edge_index"] and b.e_id.cpu().tolist() == c["e_id"]
            assert b.counts.cpu().tolist() == c["nodes_after"] + c["edges_after"]
            assert torch.equal(b.y.cpu(), torch.tensor(c["n_id"]))
