"""Mini-batch path: drop-in for `torch_geometric.loader.NeighborLoader` as the reference builds
it (`src/train_gnn.py:329-348`) -- every batch is sampled, relabelled and sliced ON THE DEVICE (`csrc/sampler.cu`).

    loader = NeighborLoader(data, num_neighbors=[10, 10], batch_size=8192, input_nodes=train_idx, shuffle=True)
    for batch in loader:
        batch = batch.to(device)                          # no-op: batches are born on the GPU
        logits = model(batch.x, batch.edge_index, batch.timestep)
        loss = loss_fn(logits[:batch.batch_size], batch.y[:batch.batch_size], ...)

`batch.x / y / timestep / *_mask` are the rows of the sampled nodes (seeds first), `batch.edge_index` the sampled edges
in local ids, `batch.n_id` / `batch.e_id` the global node ids / original edge columns, `batch.input_id` the positions of
the seeds in `input_nodes`, `batch.batch_size` the number of seeds.  One host synchronisation per batch (the two output
lengths).  PyG samples with std::mt19937 on the CPU; here a batch is a pure function of (graph, seeds, fan-outs, seed,
batch index) through a counter-based Philox stream, reproduced bit for bit by `oracle/neighbor_sample_np.py`."""
import ctypes as C
from typing import Iterator, List, Optional, Sequence

import torch

from ._lib import check, lib, ptr, stream
from .graph import build_graph

_ROW_KEYS = ("x", "y", "timestep", "train_mask", "val_mask", "test_mask")


class Batch:
    """The attributes of a PyG mini-batch `Data` that `train_epoch_minibatch` (`src/train_gnn.py:212-245`) reads."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device, *args, **kwargs):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("egnn_b200 batches live on the GPU (no CPU path)")
        return self

    @property
    def num_nodes(self) -> int:
        return int(self.n_id.numel())

    def __repr__(self):
        return (f"Batch(num_nodes={self.num_nodes}, num_edges={int(self.edge_index.size(1))}, "
                f"batch_size={self.batch_size})")


def gather_rows(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """t[idx] for a row-major tensor through `egnn_gather_rows` (one kernel, 16-byte vectors where the rows allow)."""
    t = t if t.is_contiguous() else t.contiguous()
    n = idx.numel()
    out = torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    row_bytes = t.element_size() * (t[0].numel() if t.dim() > 1 else 1)
    if n and row_bytes:
        check(lib().egnn_gather_rows(ptr(t), row_bytes, ptr(idx), n, row_bytes, ptr(out), row_bytes, stream()))
    return out


class NeighborLoader:
    def __init__(self, data, num_neighbors: Sequence[int], batch_size: int = 1, input_nodes=None, shuffle: bool = False,
                 seed: int = 0, drop_last: bool = False, device="cuda", **unsupported):
        for k, v in unsupported.items():      # PyG options the reference never passes (replace, disjoint, subgraph_type ...)
            if v not in (None, False, 0):
                raise NotImplementedError(f"NeighborLoader option {k}={v!r} is not supported")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("egnn_b200.NeighborLoader samples on the GPU (no CPU path)")
        self.fanouts: List[int] = [int(k) for k in num_neighbors]
        if not 1 <= len(self.fanouts) <= 8:
            raise ValueError("num_neighbors: 1 to 8 hops")
        self.batch_size, self.shuffle, self.drop_last, self.seed = int(batch_size), bool(shuffle), bool(drop_last), int(seed)
        self.rows = {k: getattr(data, k).to(dev).contiguous() for k in _ROW_KEYS if getattr(data, k, None) is not None}
        ei = data.edge_index.to(dev)
        self.N = int(self.rows["x"].size(0))
        self.E = int(ei.size(1))
        self.g = build_graph(ei, self.N)                 # CSR by destination = PyG's CSC (stable), + edge ids
        if input_nodes is None:
            input_nodes = torch.arange(self.N, device=dev)
        input_nodes = torch.as_tensor(input_nodes).to(dev)
        if input_nodes.dtype == torch.bool:
            input_nodes = torch.nonzero(input_nodes, as_tuple=False).view(-1)
        self.input_nodes = input_nodes.long().contiguous()
        if self.input_nodes.numel() and int(torch.unique(self.input_nodes).numel()) != self.input_nodes.numel():
            raise ValueError("input_nodes must be distinct")
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(self.seed)
        L = lib()
        B = max(1, min(self.batch_size, self.N))
        self._fan = (C.c_int32 * len(self.fanouts))(*self.fanouts)
        cn, ce = C.c_int64(0), C.c_int64(0)
        check(L.egnn_neighbor_sample_caps(self.N, self.E, B, self._fan, len(self.fanouts), C.byref(cn), C.byref(ce)))
        self.cap_nodes, self.cap_edges = int(cn.value), int(ce.value)
        i32 = dict(dtype=torch.int32, device=dev)
        self.state = torch.empty(2 * self.N, **i32)
        check(L.egnn_neighbor_sample_state_init(ptr(self.state), self.N, stream()))
        self.ws = torch.empty(L.egnn_neighbor_sample_workspace_bytes(self.N, self.E, B, self._fan, len(self.fanouts)),
                              dtype=torch.uint8, device=dev)
        self.counts = torch.zeros(2 * len(self.fanouts) + 2, **i32)
        self.info = torch.zeros(2, **i32)
        self.batches_drawn = 0
        self.dev = dev

    def __len__(self) -> int:
        n = self.input_nodes.numel()
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def sample(self, seeds: torch.Tensor, batch_idx: int, input_id: Optional[torch.Tensor] = None) -> Batch:
        """One batch for the given DISTINCT seed nodes (int64, on the device)."""
        L = lib()
        B = int(seeds.numel())
        if B < 1 or B > max(1, min(self.batch_size, self.N)):
            raise ValueError("1 <= number of seeds <= batch_size")
        i64 = dict(dtype=torch.int64, device=self.dev)
        n_id = torch.empty(self.cap_nodes, **i64)
        ei = torch.empty((2, self.cap_edges), **i64)
        e_id = torch.empty(self.cap_edges, **i64)
        check(L.egnn_neighbor_sample(ptr(self.g.csr_ptr), ptr(self.g.csr_src), ptr(self.g.csr_eid), self.N, self.E,
                                     ptr(seeds), B, self._fan, len(self.fanouts), self.seed, int(batch_idx),
                                     ptr(self.state), ptr(n_id), self.cap_nodes, ptr(ei), ptr(e_id), self.cap_edges,
                                     ptr(self.counts), ptr(self.info), ptr(self.ws), self.ws.numel(), stream()))
        n, e = (int(v) for v in self.info.tolist())          # the one synchronisation of a batch
        n_id = n_id[:n]
        fields = {k: gather_rows(v, n_id) for k, v in self.rows.items()}
        return Batch(n_id=n_id, edge_index=ei[:, :e].contiguous(), e_id=e_id[:e], batch_size=B,
                     input_id=input_id, counts=self.counts.clone(), **fields)

    def __iter__(self) -> Iterator[Batch]:
        n = self.input_nodes.numel()
        order = (torch.randperm(n, generator=self.gen, device=self.dev) if self.shuffle
                 else torch.arange(n, device=self.dev))
        for b in range(len(self)):
            pos = order[b * self.batch_size:(b + 1) * self.batch_size]
            seeds = self.input_nodes[pos].contiguous()
            yield self.sample(seeds, self.batches_drawn, input_id=pos)
            self.batches_drawn += 1
