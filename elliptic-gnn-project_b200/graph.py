"""Device-side graph structure (K1) and its cache.

The reference hands a COO int64 `edge_index` to every conv call and PyG re-derives
self-loops / degree / gcn_norm inside each GCNConv/GATConv forward
(`/root/reference/src/models/gnn.py:20-23,28,31` -> PyG gcn_norm, cached=False).  Here the
sorted views are built once per distinct `edge_index` by `egnn_graph_build` and cached,
keyed on the tensor's storage pointer, shape and in-place version counter, so that the
re-entrant callers (`src/train_gnn.py:526-540` hub ablation, `src/analysis/robustness.py:65-82`
edge dropping) that pass *different* edge tensors to the same model get their own structure.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check, lib, ptr, stream


@dataclass
class Graph:
    """Stable destination-sorted (CSR) and source-sorted (CSC) views of one edge list."""
    n_nodes: int
    cap: int                       # capacity of the per-edge arrays (upper bound on E2)
    info: torch.Tensor             # int32[4]: E2, #bad ids, len(csr_long), len(csc_long)
    csr_ptr: torch.Tensor          # int32[N+1]  rows = destinations
    csr_src: torch.Tensor          # int32[cap]
    csr_eid: torch.Tensor          # int32[cap]
    csc_ptr: torch.Tensor          # int32[N+1]  rows = sources
    csc_dst: torch.Tensor          # int32[cap]
    csc_pos: torch.Tensor          # int32[cap]  position in CSR order
    csr_long: torch.Tensor         # int32[cap/64+1]
    csc_long: torch.Tensor
    csr_order: Optional[torch.Tensor] = None   # int32[N] rows by descending degree: only the per-row fall-back
    csc_order: Optional[torch.Tensor] = None   # kernel (spmm_lean) reads it; built on request (`want_order`)
    csr_part: Optional[torch.Tensor] = None  # int32[n_tasks+1] cost-balanced row partition (streaming SpMM)
    csc_part: Optional[torch.Tensor] = None
    n_tasks: int = 0
    ei2: Optional[torch.Tensor] = None     # int64[2,cap] expanded edge list (debug/parity)
    dis: Optional[torch.Tensor] = None     # float[N]   deg^-1/2
    w_edge: Optional[torch.Tensor] = None  # float[cap] gcn_norm in ei2 order
    w_csr: Optional[torch.Tensor] = None
    w_csc: Optional[torch.Tensor] = None

    @property
    def n_edges(self) -> int:
        """E2 (host sync)."""
        return int(self.info[0].item())

    def n_long(self, view: int):
        return self.info[2 + view:3 + view]


def symmetrize(edge_index: torch.Tensor) -> torch.Tensor:
    """`cat([ei, ei.flip(0)], dim=1)` -- `/root/reference/src/train_gnn.py:321-324`; bit-exact,
    produced by the graph-build kernel's own expansion (ei2 of a SYMMETRIZE build)."""
    n = int(edge_index.max().item()) + 1 if edge_index.numel() else 1
    g = build_graph(edge_index, n, symmetrize=True, keep_edge_list=True)
    return g.ei2[:, : g.n_edges].contiguous()


def build_graph(edge_index: torch.Tensor, num_nodes: int, symmetrize: bool = False,
                self_loops: bool = False, want_norm: bool = False, keep_edge_list: bool = False,
                validate: bool = True, out: Optional[Graph] = None, want_order: bool = False) -> Graph:
    """Build the sorted views.  `out` re-uses the buffers of a previously built Graph of the same
    shape (so a captured CUDA graph that reads them sees the new structure)."""
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must have shape [2, E]")
    if edge_index.dtype != torch.int64:
        raise TypeError("edge_index must be int64")
    if not edge_index.is_cuda:
        raise RuntimeError("egnn_b200 builds graphs on the GPU only (no CPU fallback)")
    ei = edge_index.contiguous()
    dev = ei.device
    E, N = int(ei.size(1)), int(num_nodes)
    flags = (_lib.G_SYMMETRIZE if symmetrize else 0) | (_lib.G_SELF_LOOPS if self_loops else 0)
    cap = max(1, (2 * E if symmetrize else E) + (N if self_loops else 0))
    i32 = dict(dtype=torch.int32, device=dev)
    norm = want_norm or self_loops
    if out is not None:
        if out.n_nodes != N or out.cap != cap or (norm and out.w_csr is None):
            raise ValueError("`out` graph has a different shape")
    g = out if out is not None else Graph(
        n_nodes=N, cap=cap, info=torch.empty(4, **i32),
        csr_ptr=torch.empty(N + 1, **i32), csr_src=torch.empty(cap, **i32), csr_eid=torch.empty(cap, **i32),
        csc_ptr=torch.empty(N + 1, **i32), csc_dst=torch.empty(cap, **i32), csc_pos=torch.empty(cap, **i32),
        csr_long=torch.empty(cap // 64 + 1, **i32), csc_long=torch.empty(cap // 64 + 1, **i32),
        csr_order=torch.empty(N, **i32) if want_order else None,
        csc_order=torch.empty(N, **i32) if want_order else None,
        ei2=torch.empty(2, cap, dtype=torch.int64, device=dev) if keep_edge_list else None,
        dis=torch.empty(N, dtype=torch.float32, device=dev) if norm else None,
        w_edge=torch.empty(cap, dtype=torch.float32, device=dev) if (norm and keep_edge_list) else None,
        w_csr=torch.empty(cap, dtype=torch.float32, device=dev) if norm else None,
        w_csc=torch.empty(cap, dtype=torch.float32, device=dev) if norm else None,
    )
    L = lib()
    if g.csr_part is None:
        g.n_tasks = int(L.egnn_spmm_partition_tasks(N, cap))
        g.csr_part = torch.empty(g.n_tasks + 1, **i32)
        g.csc_part = torch.empty(g.n_tasks + 1, **i32)
    ws_bytes = L.egnn_graph_workspace_bytes(N, E, flags)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(L.egnn_graph_build(ptr(ei), E, N, flags, int(want_norm), ptr(g.info), ptr(g.csr_ptr), ptr(g.csr_src),
                             ptr(g.csr_eid), ptr(g.csc_ptr), ptr(g.csc_dst), ptr(g.csc_pos), ptr(g.csr_long),
                             ptr(g.csc_long), ptr(g.csr_order), ptr(g.csc_order), ptr(g.ei2), ptr(g.dis), ptr(g.w_edge), ptr(g.w_csr),
                             ptr(g.w_csc), ptr(ws), ws_bytes, stream()))
    # row partitions of both views for the streaming aggregation kernel (equal work per lane group)
    check(L.egnn_spmm_partition(ptr(g.csr_ptr), N, ptr(g.csr_part), g.n_tasks, stream()))
    check(L.egnn_spmm_partition(ptr(g.csc_ptr), N, ptr(g.csc_part), g.n_tasks, stream()))
    if validate and not torch.cuda.is_current_stream_capturing():
        bad = int(g.info[1].item())
        if bad:
            raise IndexError(f"edge_index holds {bad} node ids outside [0, {N})")
    return g


class GraphCache:
    """Keyed cache of built graphs: (data_ptr, shape, _version, num_nodes, flags)."""

    def __init__(self, max_entries: int = 8):
        self._d: Dict[Tuple, Tuple[torch.Tensor, Graph]] = {}
        self.max_entries = max_entries
        self.hits = 0
        self.misses = 0

    def get(self, edge_index: torch.Tensor, num_nodes: int, self_loops: bool = False) -> Graph:
        key = (edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride()),
               edge_index._version, int(num_nodes), bool(self_loops))
        hit = self._d.get(key)
        if hit is not None:
            self.hits += 1
            return hit[1]
        self.misses += 1
        g = build_graph(edge_index, num_nodes, self_loops=self_loops)
        if len(self._d) >= self.max_entries:
            self._d.pop(next(iter(self._d)))
        # keep a reference to the key tensor so its storage pointer cannot be recycled
        self._d[key] = (edge_index, g)
        return g

    def put(self, edge_index: torch.Tensor, num_nodes: int, self_loops: bool, g: Graph) -> None:
        """Register an already built graph for `edge_index` in its CURRENT state (train.HostFeed rebuilds into
        fixed buffers after an in-place `copy_`, which bumps `_version`)."""
        key = (edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride()),
               edge_index._version, int(num_nodes), bool(self_loops))
        for k in [k for k, v in self._d.items() if v[1] is g and k != key]:
            del self._d[k]
        self._d[key] = (edge_index, g)

    def clear(self):
        self._d.clear()


_GLOBAL_CACHE = GraphCache()


def cached_graph(edge_index: torch.Tensor, num_nodes: int, self_loops: bool = False) -> Graph:
    return _GLOBAL_CACHE.get(edge_index, num_nodes, self_loops)


def register_graph(edge_index: torch.Tensor, num_nodes: int, g: Graph, self_loops: bool = False) -> None:
    _GLOBAL_CACHE.put(edge_index, num_nodes, self_loops, g)


def ablate_hubs(edge_index: torch.Tensor, num_nodes: int, frac: float):
    """`build_edge_index_ablated` (`/root/reference/src/analysis/hub_ablation.py:56-71`, `src/train_gnn.py:526-540`) on
    the device: remove every edge touching one of the `int(frac * num_nodes)` highest-degree nodes (degree = out + in
    over the given edge list), keeping the original edge order.  The edge list never leaves the GPU (the reference
    moves it to the CPU and back); the host reads back ONE int, the number of edges kept, to size the result.
    Returns (edge_index_ablated [2, E_kept] int64, num_hubs, hub_mask bool [N]).  Ties at the k-th degree: lower node
    id first (see `egnn_hub_ablation`)."""
    if edge_index.dim() != 2 or edge_index.size(0) != 2 or edge_index.dtype != torch.int64:
        raise TypeError("edge_index must be int64 [2, E]")
    if not edge_index.is_cuda:
        raise RuntimeError("egnn_b200 ablates hubs on the GPU only (no CPU fallback)")
    ei = edge_index.contiguous()
    dev = ei.device
    E, N = int(ei.size(1)), int(num_nodes)
    num_hubs = int(float(frac) * float(N))
    L = lib()
    out = torch.empty((2, max(E, 1)), dtype=torch.int64, device=dev)
    info = torch.empty(2, dtype=torch.int32, device=dev)
    hub = torch.empty(N, dtype=torch.uint8, device=dev)
    ws_bytes = L.egnn_hub_ablation_workspace_bytes(N, E)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(L.egnn_hub_ablation(ptr(ei), E, N, num_hubs, ptr(out), ptr(info), ptr(hub), None, ptr(ws), ws_bytes, stream()))
    kept, bad = (int(v) for v in info.tolist())
    if bad:
        raise IndexError(f"edge_index holds {bad} edges with node ids outside [0, {N})")
    # rows of `out` are E apart: the kept columns as a [2, kept] tensor (contiguous copy of 16 * kept bytes)
    return out[:, :kept].contiguous() if E > 0 else out[:, :0], num_hubs, hub.bool()


def drop_edges(edge_index: torch.Tensor, drop_frac: float, perm: Optional[torch.Tensor] = None):
    """`drop_edges` (`/root/reference/src/analysis/robustness.py:65-82`): `edge_index[:, perm[drop_count:]]` with
    `perm = torch.randperm(E, device=...)` unless given (bit-exact for the same permutation).  Same argument checks and
    return convention as the reference: (edge_index_kept, drop_count)."""
    drop_frac = float(drop_frac)
    if drop_frac < 0 or drop_frac > 1:
        raise ValueError("drop_frac must be within [0, 1]")
    if drop_frac <= 0:
        return edge_index, 0
    if not edge_index.is_cuda:
        raise RuntimeError("egnn_b200 drops edges on the GPU only (no CPU fallback)")
    E = int(edge_index.size(1))
    drop_count = min(int(round(drop_frac * float(E))), E)
    if drop_count == 0:
        return edge_index, 0
    if drop_count >= E:
        raise RuntimeError("Dropping all edges would leave an empty graph.")
    if perm is None:
        perm = torch.randperm(E, device=edge_index.device)
    keep_idx = perm[drop_count:].contiguous()
    ei = edge_index.contiguous()
    out = torch.empty((2, keep_idx.numel()), dtype=torch.int64, device=ei.device)
    bad = torch.empty(1, dtype=torch.int32, device=ei.device)
    check(lib().egnn_edge_gather(ptr(ei), E, ptr(keep_idx), keep_idx.numel(), ptr(out), ptr(bad), stream()))
    return out, drop_count
