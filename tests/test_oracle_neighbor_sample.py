"""CPU checks of the neighbour-sampling restatement (oracle/neighbor_sample_np.py): hand-derived known answers on the
reference's only graph fixture (the 5-node path graph of `/root/reference/tests/test_masks_and_metrics.py:9-14`), the
structural properties PyG's NeighborLoader guarantees, and the uniformity of the Floyd draw."""
import numpy as np

from oracle.neighbor_sample_np import csc_by_destination, neighbor_sample

PATH = np.array([[0, 1, 2, 3], [1, 2, 3, 4]])
SYM = np.concatenate([PATH, PATH[::-1]], axis=1)       # src/train_gnn.py:321-324


def test_known_answers_full_fanout():
    ip, src, eid = csc_by_destination(SYM, 5)
    assert ip.tolist() == [0, 1, 3, 5, 7, 8] and src.tolist() == [1, 0, 2, 1, 3, 2, 4, 3] and eid.tolist() == [4, 0, 5, 1, 6, 2, 7, 3]
    n_id, ei, e_id, nodes, edges = neighbor_sample(ip, src, eid, [2], [-1, -1])
    # hop 1: in-neighbours of 2 = {1 (edge 1), 3 (edge 6)}; hop 2: of 1 = {0, 2}, of 3 = {2, 4}
    assert n_id.tolist() == [2, 1, 3, 0, 4]
    assert ei.tolist() == [[1, 2, 3, 0, 0, 4], [0, 0, 1, 1, 2, 2]]
    assert e_id.tolist() == [1, 6, 0, 5, 2, 7] and nodes == [1, 3, 5] and edges == [0, 2, 6]
    # directed path graph: node 0 has no in-edge -> a seed that stays alone
    ip, src, eid = csc_by_destination(PATH, 5)
    n_id, ei, e_id, nodes, edges = neighbor_sample(ip, src, eid, [0, 4], [3, 3])
    assert n_id.tolist() == [0, 4, 3, 2] and ei.tolist() == [[2, 3], [1, 2]] and e_id.tolist() == [3, 2]


def _random_graph(n, e, seed):
    rng = np.random.default_rng(seed)
    ei = rng.integers(0, n, size=(2, e))
    ei[:, : e // 10] = ei[:, e // 10: 2 * (e // 10)]            # duplicated edges
    ei[1, : e // 20] = 7                                        # a hub
    return ei


def test_structural_properties():
    n, e = 300, 2500
    ei = _random_graph(n, e, 0)
    ip, src, eid = csc_by_destination(ei, n)
    seeds = np.random.default_rng(1).permutation(n)[:40]
    for fan in ([5, 3], [2, 2, 2], [-1, 4], [1000, 1]):
        n_id, le, e_id, nodes, edges = neighbor_sample(ip, src, eid, seeds, fan, seed=11, batch_idx=3)
        assert n_id[:40].tolist() == seeds.tolist() and len(set(n_id.tolist())) == len(n_id)
        assert np.array_equal(ei[0][e_id], n_id[le[0]]) and np.array_equal(ei[1][e_id], n_id[le[1]])   # real edges
        assert len(set(e_id.tolist())) == len(e_id)                              # without replacement
        fb = 0
        for h, k in enumerate(fan):
            lo, hi = edges[h], edges[h + 1]
            dst = le[1][lo:hi]
            assert dst.min(initial=fb) >= fb and dst.max(initial=fb) < nodes[h]  # only the previous hop's nodes receive
            for vl in range(fb, nodes[h]):
                d = ip[n_id[vl] + 1] - ip[n_id[vl]]
                assert (dst == vl).sum() == (d if (k < 0 or d <= k) else k)
            assert le[0][lo:hi].max(initial=0) < nodes[h + 1]
            fb = nodes[h]


def test_floyd_draw_is_uniform():
    # node 0 with 6 in-neighbours, fan-out 2: every neighbour is picked with probability 1/3
    ei = np.array([[1, 2, 3, 4, 5, 6], [0, 0, 0, 0, 0, 0]])
    ip, src, eid = csc_by_destination(ei, 7)
    hits = np.zeros(7)
    trials = 3000
    for b in range(trials):
        n_id, le, e_id, _, _ = neighbor_sample(ip, src, eid, [0], [2], seed=99, batch_idx=b)
        assert len(e_id) == 2 and e_id[0] != e_id[1]
        hits[n_id[le[0]]] += 1
    p = hits[1:] / trials
    assert np.all(np.abs(p - 1 / 3) < 0.035), p          # 4 sigma of a Binomial(3000, 1/3) proportion


def test_oracle_reproduces_the_committed_golden_vectors():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "neighbor_sample_golden.json")
    gold = json.load(open(path))
    for name, g in gold.items():
        ei = np.asarray(g["edge_index"])
        ip, src, eid = csc_by_destination(ei, g["num_nodes"])
        for c in g["cases"]:
            n_id, le, e_id, nodes, edges = neighbor_sample(ip, src, eid, g["seeds"], c["fanouts"], seed=c["seed"],
                                                           batch_idx=c["batch_idx"])
            assert n_id.tolist() == c["n_id"] and le.tolist() == c["edge_index"] and e_id.tolist() == c["e_id"], name
            assert nodes == c["nodes_after"] and edges == c["edges_after"]
