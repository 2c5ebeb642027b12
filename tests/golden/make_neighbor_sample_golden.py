"""Golden vectors of the mini-batch neighbour sampler: `oracle/neighbor_sample_np.py` (the sequential restatement of
PyG's NeighborLoader sampling, sharing the CUDA path's Philox stream) run on two seeded graphs.  The reference itself
cannot produce them (torch_geometric / pyg-lib are not installable and draw from std::mt19937), so these
pin the (oracle, CUDA kernel) PAIR: a change of the Philox keying, of the Floyd draw or of the relabelling order in
either shows up against the committed file.  Usage: python tests/golden/make_neighbor_sample_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.neighbor_sample_np import csc_by_destination, neighbor_sample  # noqa: E402


def graphs():
    rng = np.random.default_rng(2024)
    n, e = 400, 3000
    ei = rng.integers(0, n, size=(2, e))
    ei[:, :200] = ei[:, 200:400]          # duplicated edges
    ei[1, :150] = 11                      # a hub
    ei[0, 2990:] = ei[1, 2990:]           # self-loops
    path = np.array([[0, 1, 2, 3, 1, 2, 3, 4], [1, 2, 3, 4, 0, 1, 2, 3]])
    return {"random400": (n, ei), "path5_sym": (5, path)}


def main():
    out = {}
    for name, (n, ei) in graphs().items():
        ip, src, eid = csc_by_destination(ei, n)
        seeds = (np.random.default_rng(7).permutation(n)[:min(48, n)]).tolist()
        cases = []
        for fan, seed, b in (([3, 2], 5, 0), ([10, 10], 42, 3), ([-1, 2], 1, 1), ([2, 2, 2], 9, 7)):
            n_id, le, e_id, nodes, edges = neighbor_sample(ip, src, eid, seeds, fan, seed=seed, batch_idx=b)
            cases.append({"fanouts": fan, "seed": seed, "batch_idx": b, "n_id": n_id.tolist(),
                          "edge_index": le.tolist(), "e_id": e_id.tolist(), "nodes_after": nodes, "edges_after": edges})
        out[name] = {"num_nodes": n, "edge_index": ei.tolist(), "seeds": seeds, "cases": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neighbor_sample_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
