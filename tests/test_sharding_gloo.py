"""Timestep-sharded data parallelism on CPU with the gloo backend, world_size 2.

What is checked without a GPU: the partition is a contiguous, halo-free split of whole timestep blocks;
the global loss normaliser / class weights come out of the all-reduce; BatchNorm statistics reduced
through `StatsReducer` equal the full-batch statistics; and -- using the CPU oracle as the per-rank
model -- the all-reduced gradient of the shard losses equals the single-process gradient on the
whole graph (the property the B200 path relies on: zero feature exchange, gradients only).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyg_restated as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _graph():
    from egnn_b200 import synthetic
    return synthetic.make_elliptic_like(n_nodes=3000, n_edges=3600, n_feats=24, n_timesteps=10, seed=5,
                                        hub_degree=60, t_train_end=7, t_val_end=8, train_window_k=5)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from egnn_b200.shard import ShardedContext, make_shard
    gr = _graph()
    sh = make_shard(gr, rank, world)
    lg = sh.graph
    ctx = ShardedContext(sh, torch.device("cpu"))
    # --- BatchNorm statistics through the reducer (sum, sumsq over the shard's rows)
    st = torch.stack([lg.x.double().sum(0), (lg.x.double() ** 2).sum(0)])
    ctx.stats_reducer.reduce_(st)
    # --- per-rank oracle step on the shard: local loss SUM / global count, gradients all-reduced
    cfg = dict(hidden_dim=16, layers=3, dropout=0.0)
    torch.manual_seed(0)
    model = O.build_model("sage", lg.x.size(1), cfg)
    ei = torch.cat([lg.edge_index, lg.edge_index.flip(0)], 1)
    logits = model(lg.x, ei, None)
    m = lg.train_mask
    per = torch.nn.functional.cross_entropy(logits[m], lg.y[m], weight=ctx.class_weight, reduction="none")
    loss = per.sum() / ctx.n_train_total
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    ctx.reduce_grads(flat)
    lsum = loss.detach().clone()
    dist.all_reduce(lsum)
    # --- epoch tail over all shards: per-row scores / labels / masks gathered with padding (ShardedContext.gather_rows)
    score = torch.softmax(logits.detach(), 1)[:, 1].contiguous()
    g_score = ctx.gather_rows(score, 0.0)
    g_y = ctx.gather_rows(lg.y, -1)
    g_mask = ctx.gather_rows(lg.val_mask.to(torch.uint8), 0)
    if rank == 0:
        torch.save({"row0": sh.row0, "n_local": sh.n_local, "n_train_total": ctx.n_train_total,
                    "cw": ctx.class_weight, "stats": st, "grad": flat, "loss": lsum,
                    "g_score": g_score, "g_y": g_y, "g_mask": g_mask}, out_path)
    bounds = torch.tensor([sh.row0, sh.row0 + sh.n_local])
    got = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(got, bounds)
    if rank == 0:
        torch.save(torch.stack(got), out_path + ".bounds")
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_the_single_process_step(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    bounds = torch.load(out + ".bounds")
    gr = _graph()
    # contiguous, disjoint, covering, cut only at timestep boundaries
    assert bounds[0, 0] == 0 and bounds[0, 1] == bounds[1, 0] and bounds[1, 1] == gr.num_nodes
    cut = int(bounds[0, 1])
    assert gr.timestep[cut - 1] != gr.timestep[cut]
    # no edge crosses the cut (zero halo)
    side = gr.edge_index >= cut
    assert torch.equal(side[0], side[1])
    # global counts and class weights
    tm = gr.train_mask
    assert r["n_train_total"] == float(tm.sum())
    assert torch.allclose(r["cw"], O.class_weight(gr.y[tm]))
    # BatchNorm statistics
    full = torch.stack([gr.x.double().sum(0), (gr.x.double() ** 2).sum(0)])
    assert torch.allclose(r["stats"], full, rtol=1e-12, atol=1e-9)
    # single-process reference step on the whole graph
    cfg = dict(hidden_dim=16, layers=3, dropout=0.0)
    torch.manual_seed(0)
    model = O.build_model("sage", gr.x.size(1), cfg)
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    loss = O.masked_weighted_ce(model(gr.x, ei, None), gr.y, tm, O.class_weight(gr.y[tm]))
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert abs(float(r["loss"]) - float(loss.detach())) <= 1e-6 * abs(float(loss.detach()))
    assert (r["grad"] - flat).abs().max() <= 1e-5 * flat.abs().max()
    # the gathered (padded, masked) rows carry exactly the whole graph's validation set: same PR-AUC as one process
    from oracle import metrics_np as M
    with torch.no_grad():
        full_score = torch.softmax(model(gr.x, ei, None), 1)[:, 1].numpy()
    vm = gr.val_mask.numpy()
    want = M.average_precision((gr.y.numpy()[vm] == 1).astype(int), full_score[vm])
    gm = r["g_mask"].numpy().astype(bool)
    got = M.average_precision((r["g_y"].numpy()[gm] == 1).astype(int), r["g_score"].numpy()[gm])
    assert got[1:] == want[1:] and abs(got[0] - want[0]) <= 1e-6


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_covers_all_units_contiguously(world):
    from egnn_b200.shard import make_shard, partition_contiguous
    gr = _graph()
    rows = [(make_shard(gr, r, world).row0, make_shard(gr, r, world).n_local) for r in range(world)]
    pos = 0
    for row0, n in rows:
        assert row0 == pos
        pos += n
    assert pos == gr.num_nodes
    parts = partition_contiguous([3, 1, 4, 1, 5, 9, 2, 6], world)
    assert parts[0][0] == 0 and parts[-1][1] == 8 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
