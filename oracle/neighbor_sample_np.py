"""TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else).

CPU restatement of the neighbour sampling behind `torch_geometric.loader.NeighborLoader` as the reference uses it
(`/root/reference/src/train_gnn.py:329-348`: `NeighborLoader(data, num_neighbors=fanout, batch_size=batch_size,
input_nodes=train_idx, shuffle=True)`; consumed by `train_epoch_minibatch`, `:212-245`).

PARITY UNPINNED: torch_geometric / pyg-lib are not installable here and the reference's tests never
build a loader.  The algorithm restated is the published one of PyG 2.5.3 (`torch_geometric/sampler/neighbor_sampler.py`
-> pyg-lib `neighbor_sample`, homogeneous, replace=False, directed=True, disjoint=False):
  * the node list starts with the seeds; for hop h with fan-out k, every node v added by the previous hop (in order) takes
    ALL in-neighbours when indeg(v) <= k or k < 0, else k distinct ones (Robert Floyd's algorithm over the positions of
    v's CSC row: for j = d-k .. d-1: t = uniform{0..j}; take t unless already taken, else j);
  * a sampled source u joins the node list at its first appearance; local id = position in the node list;
  * edge (u -> v) is emitted as (local(u), local(v)) in sampling order.
pyg-lib draws from std::mt19937 -- a stream this restatement does not reproduce; the CUDA path and this file share a
counter-based stream instead (Philox4x32-10 keyed on (seed, batch index, hop, local id of v, draw): csrc/sampler.cu), so
the two agree bit for bit and the distributional / structural properties are what pins them to PyG.
"""
import numpy as np

from .graph_build_np import philox4x32_10


def csc_by_destination(edge_index: np.ndarray, n_nodes: int):
    """indptr [N+1], src [E], eid [E]: in-edges of every node in their ORIGINAL order (stable sort by destination)."""
    dst = edge_index[1]
    order = np.argsort(dst, kind="stable")
    indptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(indptr, dst + 1, 1)
    return np.cumsum(indptr), edge_index[0][order], order


def _draw(vl: int, j: int, h: int, batch_idx: int, seed: int) -> int:
    w = philox4x32_10(np.uint32(vl), np.uint32(j >> 2), np.uint32(h), np.uint32(batch_idx & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return int(np.asarray(w[j & 3]).reshape(-1)[0])


def neighbor_sample(indptr, src, eid, seeds, fanouts, seed: int = 0, batch_idx: int = 0):
    """Sequential restatement.  Returns n_id [n], edge_index [2, e] (local ids), e_id [e], nodes_after [H+1],
    edges_after [H+1]."""
    local, nodes = {}, []
    for s in seeds:
        local[int(s)] = len(nodes)
        nodes.append(int(s))
    e_src, e_dst, e_ids = [], [], []
    fb, fe = 0, len(nodes)
    nodes_after, edges_after = [len(nodes)], [0]
    for h, k in enumerate(fanouts):
        for vl in range(fb, fe):
            v = nodes[vl]
            row, d = int(indptr[v]), int(indptr[v + 1] - indptr[v])
            if k < 0 or d <= k:
                pos = list(range(d))
            else:
                pos = []
                for j in range(k):
                    jj = d - k + j
                    t = (_draw(vl, j, h, batch_idx, seed) * (jj + 1)) >> 32
                    pos.append(jj if t in pos else t)
            for p in pos:
                u = int(src[row + p])
                if u not in local:
                    local[u] = len(nodes)
                    nodes.append(u)
                e_src.append(local[u])
                e_dst.append(vl)
                e_ids.append(int(eid[row + p]))
        fb, fe = fe, len(nodes)
        nodes_after.append(len(nodes))
        edges_after.append(len(e_src))
    return (np.asarray(nodes, dtype=np.int64), np.asarray([e_src, e_dst], dtype=np.int64).reshape(2, -1),
            np.asarray(e_ids, dtype=np.int64), nodes_after, edges_after)
