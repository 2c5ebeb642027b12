"""Timestep-sharded data parallelism.

The reference loader drops every cross-timestep edge (`/root/reference/src/data/
dataset_elliptic.py:235-241`; invariant stated by `src/analysis/eda.py:124-150`), so the graph
is a disjoint union of per-timestep blocks and node order is timestep-contiguous.  Whole
(replica, timestep) blocks are assigned to ranks as CONTIGUOUS ranges (so a shard's global
node ids are one interval and the Philox dropout mask keyed on `row0 + local_row` equals the
single-GPU mask); every rank holds its rows of x / y / masks / timestep and a local edge list
with re-based ids: zero halo, zero feature exchange.  Collectives (NCCL over NVLink):
  * one all-reduce of the flat weight-gradient buffer per step;
  * per BatchNorm layer one all-reduce of [sum, sumsq] (fwd) and [sum g, sum g*xhat] (bwd),
    2 x hidden doubles each;
  * once per run: global train-row count and class counts (loss normaliser, class weights).
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .ops import StatsReducer
from .synthetic import EllipticGraph


def partition_contiguous(weights: Sequence[float], n_ranks: int) -> List[Tuple[int, int]]:
    """Split `weights` into n_ranks contiguous ranges [lo, hi) minimising the maximum range sum
    (binary search on the bottleneck + greedy fill).  Empty ranges are allowed only when there
    are fewer units than ranks."""
    w = [float(v) for v in weights]
    n = len(w)
    if n_ranks <= 0:
        raise ValueError("n_ranks must be positive")

    def fits(cap: float) -> Optional[List[Tuple[int, int]]]:
        out, lo, acc = [], 0, 0.0
        for i, v in enumerate(w):
            if v > cap:
                return None
            if acc + v > cap:
                out.append((lo, i))
                lo, acc = i, 0.0
            acc += v
        out.append((lo, n))
        return out if len(out) <= n_ranks else None

    lo_c, hi_c = max(w, default=0.0), sum(w)
    best = fits(hi_c)
    for _ in range(60):
        mid = 0.5 * (lo_c + hi_c)
        r = fits(mid)
        if r is None:
            lo_c = mid
        else:
            best, hi_c = r, mid
    # spread trailing empties: split the largest multi-unit ranges until we have n_ranks pieces
    while len(best) < n_ranks:
        cand = [(sum(w[a:b]), k) for k, (a, b) in enumerate(best) if b - a >= 2]
        if not cand:
            break
        _, k = max(cand)
        a, b = best[k]
        run, cut, half = 0.0, a + 1, 0.5 * sum(w[a:b])
        for i in range(a, b - 1):
            run += w[i]
            cut = i + 1
            if run >= half:
                break
        best[k:k + 1] = [(a, cut), (cut, b)]
    while len(best) < n_ranks:
        best.append((n, n))
    return best


@dataclass
class Shard:
    rank: int
    world: int
    row0: int            # global id of this shard's first node
    n_local: int
    n_global: int
    graph: EllipticGraph  # local rows, edge ids re-based to [0, n_local)


def unit_bounds(timestep: torch.Tensor) -> torch.Tensor:
    """Start offsets of the maximal runs of equal timestep value (= (replica, timestep) blocks);
    last entry = N."""
    n = timestep.numel()
    change = torch.nonzero(timestep[1:] != timestep[:-1]).view(-1) + 1
    return torch.cat([torch.zeros(1, dtype=torch.int64), change, torch.tensor([n])])


def make_shard(gr: EllipticGraph, rank: int, world: int, feat_cost: float = 1.0) -> Shard:
    """Cut the rank's contiguous block range out of a (CPU) graph.  Load model per block:
    nodes*F (feature traffic) + edges (index traffic)."""
    bounds = unit_bounds(gr.timestep)
    n_units = bounds.numel() - 1
    ei = gr.edge_index
    unit_of_node = torch.bucketize(torch.arange(gr.num_nodes), bounds[1:], right=True)
    e_unit = unit_of_node[ei[1]]
    if not torch.equal(e_unit, unit_of_node[ei[0]]):
        raise ValueError("edge crosses a timestep block: the graph does not shard without halo")
    e_per_unit = torch.bincount(e_unit, minlength=n_units)
    n_per_unit = bounds[1:] - bounds[:-1]
    w = (n_per_unit.double() * feat_cost * gr.x.size(1) + e_per_unit.double()).tolist()
    lo_u, hi_u = partition_contiguous(w, world)[rank]
    lo = int(bounds[lo_u]) if lo_u < n_units else gr.num_nodes
    hi = int(bounds[hi_u]) if hi_u <= n_units else gr.num_nodes
    keep = (ei[1] >= lo) & (ei[1] < hi)
    sl = lambda t: None if t is None else t[lo:hi].contiguous()
    local = EllipticGraph(x=gr.x[lo:hi].contiguous(), edge_index=(ei[:, keep] - lo).contiguous(),
                          y=sl(gr.y), timestep=sl(gr.timestep), train_mask=sl(gr.train_mask),
                          val_mask=sl(gr.val_mask), test_mask=sl(gr.test_mask))
    return Shard(rank=rank, world=world, row0=lo, n_local=hi - lo, n_global=gr.num_nodes, graph=local)


class P2PAllReduce:
    """In-place sum over the ranks of a small fp32 / fp64 vector as ONE hand-written kernel over NVLink peer
    memory (csrc/p2p.cu: push into the peers' symmetric buffers, per-chunk flags, sum in rank order) instead
    of a NCCL call.  The symmetric buffer and the device array of peer pointers come from
    torch.distributed._symmetric_memory (allocation + rendezvous only)."""

    def __init__(self, n_max: int, dtype: torch.dtype, group, device):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.code = {torch.float32: _lib.F32, torch.float64: _lib.F64}[dtype]
        self.dtype, self.n_max, self.group = dtype, int(n_max), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        nbytes = _lib.lib().egnn_p2p_allreduce_buffer_bytes(self.world, self.n_max, self.code)
        self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.peer_ptrs = int(self.hdl.buffer_ptrs_dev)
        self.epoch = torch.zeros(2, dtype=torch.int64, device=device)   # {calls completed, ticket of the running call}
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.timeout_ms = int(os.environ.get("EGNN_P2P_TIMEOUT_MS", "2000"))
        torch.cuda.synchronize(device)   # flags are zero; the caller's consensus all-reduce orders this before any push

    @staticmethod
    def create_on_all_ranks(n_max: int, dtype: torch.dtype, group, device):
        """Build the reducer on every rank or on none: a rank whose symmetric-memory set-up failed must not leave
        the others waiting on its flags, so the outcome is agreed on with one NCCL MIN all-reduce (which also
        orders every rank's zero-initialisation before the first push)."""
        ar, why = None, ""
        try:
            ar = P2PAllReduce(n_max, dtype, group, device)
        except Exception as ex:      # no symmetric-memory support on this box / driver
            why = f"{type(ex).__name__}: {ex}"
        ok = torch.tensor([1 if ar is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if dist.get_rank(group) == 0:
                print(f"[egnn_b200] peer-memory all-reduce unavailable ({why or 'failed on another rank'}); using NCCL",
                      file=sys.stderr, flush=True)
            return None
        return ar

    def supports(self, t: torch.Tensor) -> bool:
        return t.dtype == self.dtype and t.is_contiguous() and 0 < t.numel() <= self.n_max

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        from ._lib import check, lib, stream
        check(lib().egnn_p2p_allreduce(t.data_ptr(), t.data_ptr(), t.numel(), self.code, self.n_max, self.peer_ptrs,
                                       self.rank, self.world, self.epoch.data_ptr(), self.err.data_ptr(),
                                       self.timeout_ms, stream()))
        return t

    def check(self):
        """Host synchronisation point: raise if any call since the start timed out waiting for a peer (the kernel
        has already poisoned its output with NaN; the flag is sticky)."""
        if int(self.err.item()):
            raise RuntimeError(f"egnn_p2p_allreduce: a peer did not arrive within {self.timeout_ms} ms "
                               "(the affected result was set to NaN)")


class P2PStatsReducer(StatsReducer):
    """BatchNorm-statistics all-reduce through the peer-memory kernel (NCCL for anything it does not cover).
    `fused_args()` hands the raw exchange arguments to the single-kernel reduce + exchange + finalise entry points
    (`egnn_bn_stats_exchange`, `egnn_bn_bwd_sums_exchange`) used by `fused.py`."""

    def __init__(self, n_total: int, group, ar: "P2PAllReduce"):
        super().__init__(n_total=n_total, group=group)
        self.ar = ar

    def fused_args(self, n_feat: int):
        """(n_max, peer_ptrs, rank, world, epoch_ptr, err_ptr, timeout_ms) or None when 2 * n_feat does not fit."""
        a = self.ar
        if a.dtype != torch.float64 or 2 * n_feat > min(a.n_max, 1024):
            return None
        return (a.n_max, a.peer_ptrs, a.rank, a.world, a.epoch.data_ptr(), a.err.data_ptr(), a.timeout_ms)

    def reduce_(self, buf: torch.Tensor) -> torch.Tensor:
        return self.ar(buf) if self.ar.supports(buf) else super().reduce_(buf)


class ShardedContext:
    """What a rank needs to step its shard so that the result equals the single-GPU step on the
    whole graph (up to the order of the cross-rank sums)."""

    def __init__(self, shard: Shard, device, group=None):
        self.shard, self.group = shard, group
        y, tm = shard.graph.y, shard.graph.train_mask
        counts = torch.tensor([float(tm.sum()), float(((y == 1) & tm).sum()), float(((y == 0) & tm).sum())],
                              dtype=torch.float64, device=device)
        if dist.is_initialized():
            dist.all_reduce(counts, group=group)
        self.n_train_total, pos, neg = (float(v) for v in counts.tolist())
        if pos == 0 or neg == 0:      # class_weight(), src/train_gnn.py:116-123, on GLOBAL counts
            self.class_weight = torch.tensor([1.0, 1.0], dtype=torch.float32)
        else:
            self.class_weight = torch.tensor([(pos + neg) / (2.0 * neg), (pos + neg) / (2.0 * pos)],
                                             dtype=torch.float32)
        multi = dist.is_initialized() and dist.get_world_size(group) > 1
        # one rank: no reducer at all, so the model takes its single-GPU fast path (statistics finalised in one launch)
        self.stats_reducer = StatsReducer(n_total=shard.n_global, group=group) if multi else None
        self.p2p = False
        self._grad_ar = None
        self._grad_ar_tried = False
        if dist.is_initialized() and dist.get_world_size(group) > 1 and torch.device(device).type == "cuda" \
                and os.environ.get("EGNN_P2P", "1") != "0":
            ar = P2PAllReduce.create_on_all_ranks(1024, torch.float64, group, device)
            if ar is not None:
                self.stats_reducer = P2PStatsReducer(shard.n_global, group, ar)
                self.p2p = True

    # ---- epoch tail over all shards -------------------------------------------------
    def gather_rows(self, t: torch.Tensor, fill=0) -> torch.Tensor:
        """Concatenation over the ranks of a per-row tensor `[n_local, ...]`, every rank's block padded with `fill`
        to the largest shard (shards are whole timesteps, so their sizes differ): `[world * n_max, ...]`, identical
        on every rank.  One all_gather; no host synchronisation."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return t
        world = dist.get_world_size(self.group)
        if not hasattr(self, "_n_max"):
            sizes = torch.zeros(world, dtype=torch.int64, device=t.device)
            sizes[dist.get_rank(self.group)] = t.size(0)
            dist.all_reduce(sizes, group=self.group)
            self._n_max = int(sizes.max().item())          # once per run
        pad = torch.full((self._n_max,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
        pad[: t.size(0)] = t
        out = torch.empty((world * self._n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather(list(out.chunk(world)), pad, group=self.group)   # views of `out`: gloo and nccl alike
        return out

    def average_precision(self, y: torch.Tensor, mask: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
        """GLOBAL validation PR-AUC (`pr_auc_illicit` over the whole graph's mask): every rank scores its own rows
        (`egnn_average_precision` with `scores_out`), the per-row scores / labels / masks are all-gathered (padding
        rows masked out) and the same device kernel ranks the union on every rank -> float64[4], identical
        everywhere.  With one rank this is `metrics.average_precision`."""
        from . import metrics
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return metrics.average_precision(y, mask, logits=logits)
        scores = torch.empty(y.numel(), dtype=torch.float32, device=y.device)
        metrics.average_precision(y, mask, logits=logits, scores_out=scores)
        m8 = mask.to(torch.uint8)
        return metrics.average_precision(self.gather_rows(y, -1), self.gather_rows(m8, 0),
                                         scores=self.gather_rows(scores, 0.0))

    def check(self):
        """Call wherever the host synchronises anyway (loss read-back, early-stopping poll): raises when a
        peer-memory all-reduce of this context timed out (`P2PAllReduce.check`)."""
        ar = getattr(self.stats_reducer, "ar", None)
        if ar is not None:
            ar.check()
        if self._grad_ar is not None:
            self._grad_ar.check()

    def attach(self, model):
        model.stats_reducer = self.stats_reducer
        model.row0 = self.shard.row0
        return model

    def reduce_grads(self, flat_grad: torch.Tensor):
        if not dist.is_initialized():
            return flat_grad
        if self.p2p and self._grad_ar is None and not self._grad_ar_tried and flat_grad.dtype == torch.float32 \
                and flat_grad.numel() <= 131072 and not torch.cuda.is_current_stream_capturing():
            self._grad_ar_tried = True     # the first (eager) step of every rank gets here together
            self._grad_ar = P2PAllReduce.create_on_all_ranks(flat_grad.numel(), torch.float32, self.group,
                                                             flat_grad.device)
        if self._grad_ar is not None and self._grad_ar.supports(flat_grad):
            return self._grad_ar(flat_grad)
        dist.all_reduce(flat_grad, group=self.group)
        return flat_grad
