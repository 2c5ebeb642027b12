"""train.HostFeed (double-buffered host -> device feed of the step inputs): same losses as stepping on resident
tensors, new host contents are picked up on the next step, the graph structure follows a changed edge list."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(hidden_dim=32, layers=3, dropout=0.0, time_embed_dim=2, time_embed_type="sin", max_timestep=49)


def _setup(egnn, capture):
    from egnn_b200 import synthetic
    from egnn_b200.train import TrainStep
    gr = synthetic.make_elliptic_like(n_nodes=5000, n_edges=6000, n_timesteps=10, seed=2, hub_degree=80,
                                      t_train_end=7, t_val_end=8)
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1).contiguous()
    host = {"x": gr.x, "ei": ei, "t": gr.timestep, "y": gr.y, "m": gr.train_mask}
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    dev = {k: v.cuda() for k, v in host.items()}
    torch.manual_seed(0)
    model = egnn.build_model("sage_resbn", 166, CFG).cuda()
    step = TrainStep(model, dev["x"], dev["ei"], dev["t"], dev["y"], dev["m"], lr=1e-3, weight_decay=0.0, amp=False)
    step.run()
    if capture:
        step.capture(warmup=1)
    return gr, host, dev, step


@pytest.mark.parametrize("ahead", [1, 2])
@pytest.mark.parametrize("capture", [False, True])
def test_feed_reproduces_resident_steps(egnn, capture, ahead):
    """`ahead` = submissions in flight: 1 = submit(); run(); submit(); ...  2 = one submission ahead (two staging sets)."""
    from egnn_b200.train import HostFeed
    gr, host, dev, step = _setup(egnn, capture)
    ref = [float(step.run()) for _ in range(5)]
    gr2, host2, dev2, step2 = _setup(egnn, capture)           # identical second run, fed from the host
    g = egnn.cached_graph(dev2["ei"], gr2.num_nodes)
    feed = HostFeed(step2, host2, dev2, gr2.num_nodes, g)
    got = []
    for _ in range(ahead):
        feed.submit()
    for i in range(5):
        prev = feed.run()
        if i + ahead < 5:
            feed.submit()
        if prev is not None:
            got.append(prev)
    got.append(feed.drain())
    assert got == ref
    assert feed.rebuilds == 0          # unchanged edge list: the sorted views are never rebuilt
    feed.submit()
    feed.submit()
    with pytest.raises(RuntimeError, match="already waiting"):
        feed.submit()


def test_feed_picks_up_new_host_contents(egnn):
    from egnn_b200.train import HostFeed
    gr, host, dev, step = _setup(egnn, False)
    g = egnn.cached_graph(dev["ei"], gr.num_nodes)
    feed = HostFeed(step, host, dev, gr.num_nodes, g)
    feed.submit()
    feed.run()
    # new features and a new edge list (half of the edges removed: same capacity, fewer entries are NOT allowed by
    # the in-place rebuild, so the edge list is permuted instead -- same multiset, different order)
    host["x"].mul_(0.5)
    perm = torch.randperm(host["ei"].size(1))
    host["ei"].copy_(host["ei"][:, perm])
    feed.submit()
    feed.run()
    feed.drain()
    assert torch.equal(dev["x"].cpu(), host["x"]) and torch.equal(dev["ei"].cpu(), host["ei"])
    assert feed.rebuilds == 1          # the permuted edge list was detected on the device, the first submit was not
    assert egnn.cached_graph(dev["ei"], gr.num_nodes) is g   # eager steps find the rebuilt graph (no second build)
    fresh = egnn.build_graph(dev["ei"], gr.num_nodes)
    assert torch.equal(g.csr_src[: g.n_edges], fresh.csr_src[: fresh.n_edges])
    assert torch.equal(g.csr_part, fresh.csr_part)


def test_feed_rejects_changed_mask_or_labels(egnn):
    """TrainStep reads the train mask / labels once; a feed that changes them fails loudly instead of
    stepping on stale train-row indices and class weights."""
    from egnn_b200.train import HostFeed
    for key in ("m", "y"):
        gr, host, dev, step = _setup(egnn, False)
        g = egnn.cached_graph(dev["ei"], gr.num_nodes)
        feed = HostFeed(step, host, dev, gr.num_nodes, g)
        assert set(feed.frozen) == {"m", "y"}
        feed.submit()
        feed.run()
        if key == "m":
            host["m"][:10] = ~host["m"][:10]
        else:
            host["y"][:10] = 1 - host["y"][:10].clamp(min=0)
        feed.submit()
        with pytest.raises(RuntimeError, match="build a new TrainStep"):
            feed.run()
