"""Synthetic Elliptic-shaped transaction graph.

The real CSVs are Git-LFS stubs (`/root/reference/data/raw/*.csv`), so every bench and
parity test runs on this seeded stand-in.  It honours the *output contract* of the
reference loader (`src/data/dataset_elliptic.py:190-196,235-245`): nodes contiguous by
timestep, `edge_index` int64 [2,E], every edge inside one timestep, `y` in {-1,0,1},
`timestep` int64 in 1..49.  Pure torch-CPU with an explicit Generator, so the CPU oracle
and the CUDA path see bit-identical inputs.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Optional

import torch

N_NODES, N_EDGES, N_FEATS, N_TIMESTEPS = 203_769, 234_355, 166, 49


@dataclass
class EllipticGraph:
    x: torch.Tensor  # [N, F] float32
    edge_index: torch.Tensor  # [2, E] int64, directed, intra-timestep
    y: torch.Tensor  # [N] int64 in {-1, 0, 1}
    timestep: torch.Tensor  # [N] int64 in 1..T
    train_mask: Optional[torch.Tensor] = None
    val_mask: Optional[torch.Tensor] = None
    test_mask: Optional[torch.Tensor] = None

    @property
    def num_nodes(self) -> int:
        return int(self.x.size(0))

    def sha256(self) -> dict:
        h = lambda t: hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]
        return {"edge_index": h(self.edge_index), "x": h(self.x)}


def _timestep_sizes(n_nodes: int, n_t: int, g: torch.Generator) -> torch.Tensor:
    lo, hi = max(1, min(1000, n_nodes // (2 * n_t))), max(2, min(8000, 4 * n_nodes // n_t))
    w = -torch.log(torch.rand(n_t, 4, generator=g)).sum(1)  # Gamma(4) draws -> Dirichlet weights
    sizes = (w / w.sum() * n_nodes).floor().long().clamp(lo, hi)
    # fix the total by round-robin +-1 inside the clip range
    diff = int(n_nodes - sizes.sum())
    i = 0
    while diff != 0:
        j = i % n_t
        step = 1 if diff > 0 else -1
        if lo <= sizes[j] + step <= hi:
            sizes[j] += step
            diff -= step
        i += 1
    return sizes


def _timestep_edges(n: int, m: int, g: torch.Generator, hub_deg: int = 0) -> torch.Tensor:
    """m distinct directed non-loop edges over local ids 0..n-1: a preferential random
    recursive tree (n-1 edges, random direction) + optional hub star + random extras, with
    ~0.5 % reciprocal pairs."""
    m = max(m, n - 1)
    child = torch.arange(1, n, dtype=torch.int64)
    u = torch.rand(n - 1, generator=g)
    parent = (child.double() * u.double().pow(2.0)).floor().long().clamp_(0, n - 1)
    parent = torch.minimum(parent, child - 1)
    flip = torch.rand(n - 1, generator=g) < 0.5
    src = torch.where(flip, child, parent)
    dst = torch.where(flip, parent, child)
    parts = [torch.stack([src, dst])]
    if hub_deg > 0:
        hub = n // 2
        others = torch.randperm(n, generator=g)[: hub_deg + 1]
        others = others[others != hub][:hub_deg]
        parts.append(torch.stack([torch.full_like(others, hub), others]))
    n_rec = max(1, int(0.005 * m))
    e0 = parts[0][:, torch.randperm(n - 1, generator=g)[:n_rec]]
    parts.append(e0.flip(0))
    e = torch.cat(parts, dim=1)
    while True:
        key = e[0] * n + e[1]
        # order-preserving dedup: keep the first occurrence of every (src,dst)
        sk, order = torch.sort(key, stable=True)
        keep_sorted = torch.ones_like(sk, dtype=torch.bool)
        keep_sorted[1:] = sk[1:] != sk[:-1]
        keep = torch.zeros_like(keep_sorted)
        keep[order] = keep_sorted
        keep &= e[0] != e[1]
        e = e[:, keep]
        if e.size(1) >= m:
            break
        need = m - e.size(1)
        extra = torch.randint(0, n, (2, need + need // 8 + 8), generator=g)
        e = torch.cat([e, extra], dim=1)
    if e.size(1) > m:
        # never drop tree / hub / reciprocal edges (they sit at the front)
        e = e[:, :m]
    return e


def make_elliptic_like(n_nodes: int = N_NODES, n_edges: int = N_EDGES, n_feats: int = N_FEATS,
                       n_timesteps: int = N_TIMESTEPS, seed: int = 42, hub_degree: int = 480,
                       t_train_end: int = 34, t_val_end: int = 43,
                       train_window_k: Optional[int] = None, label_signal: float = 0.0) -> EllipticGraph:
    """`label_signal = 0` (default, the bench workload): labels independent of features, as SURVEY.md Appendix C
    specifies.  `label_signal > 0`: the illicit class is PLANTED -- a node is illicit when a fixed random linear score
    of its own features plus `label_signal` times the mean score of its neighbours (plus unit noise) is in the top
    9.65 % of the labelled nodes -- so that a trained model reaches a PR-AUC well above chance and the 3-decimal
    PR-AUC / F1 comparison of the trajectory tests is a statement about the model, not about tie-breaking noise."""
    g = torch.Generator().manual_seed(seed)
    sizes = _timestep_sizes(n_nodes, n_timesteps, g)
    # edges per timestep proportional to nodes, exact total
    m = (sizes.double() / n_nodes * n_edges).floor().long()
    m = torch.maximum(m, sizes - 1)
    diff = int(n_edges - m.sum())
    i = 0
    while diff != 0:
        j = i % n_timesteps
        step = 1 if diff > 0 else -1
        if m[j] + step >= sizes[j] - 1:
            m[j] += step
            diff -= step
        i += 1
    hub_t = int(torch.argmax(sizes))
    hub_degree = min(hub_degree, int(sizes[hub_t]) // 2)
    blocks, offs = [], 0
    for t in range(n_timesteps):
        e = _timestep_edges(int(sizes[t]), int(m[t]), g, hub_degree if t == hub_t else 0)
        blocks.append(e + offs)
        offs += int(sizes[t])
    ei = torch.cat(blocks, dim=1)
    ei = ei[:, torch.randperm(ei.size(1), generator=g)].contiguous()
    assert ei.size(1) == n_edges, (ei.size(1), n_edges)
    timestep = torch.repeat_interleave(torch.arange(1, n_timesteps + 1), sizes)
    x = _features(n_nodes, n_feats, g)
    r = torch.rand(n_nodes, generator=g)
    y = torch.full((n_nodes,), -1, dtype=torch.int64)
    y[r < 0.228] = 0
    y[r < 0.022] = 1
    if label_signal > 0:
        w = torch.randn(n_feats, generator=g) / float(n_feats) ** 0.5
        s = x.clamp(-4, 4) @ w
        src, dst = torch.cat([ei[0], ei[1]]), torch.cat([ei[1], ei[0]])
        nb = torch.zeros(n_nodes).index_add_(0, dst, s[src])
        nb = nb / torch.bincount(dst, minlength=n_nodes).clamp(min=1).float()
        score = s + float(label_signal) * nb + torch.randn(n_nodes, generator=g)
        labeled = r < 0.228
        k = max(1, int(round(0.0965 * int(labeled.sum()))))
        thr = torch.topk(score[labeled], k).values[-1]
        y = torch.where(labeled, (score >= thr).long(), torch.full_like(y, -1))
    out = EllipticGraph(x=x, edge_index=ei, y=y, timestep=timestep)
    set_temporal_masks(out, t_train_end, t_val_end, train_window_k)
    return out


def _features(n: int, f: int, g: torch.Generator) -> torch.Tensor:
    """Drawn on the generator's device (CPU for everything the oracle sees; bench.py draws the replicas of the 64x
    scale-up directly on the GPU)."""
    dev = g.device
    x = torch.randn(n, f, generator=g, device=dev)
    n_heavy = max(1, f // 20)  # ~5 % heavy-tailed columns (standardised features with outliers)
    cols = torch.randperm(f, generator=g, device=dev)[:n_heavy]
    x[:, cols] = x[:, cols] * torch.exp(0.75 * torch.randn(n, n_heavy, generator=g, device=dev))
    return x


def set_temporal_masks(gr: EllipticGraph, t_train_end: int, t_val_end: int,
                       train_window_k: Optional[int] = None) -> EllipticGraph:
    """`make_temporal_masks` (`src/data/dataset_elliptic.py:268-290`) on an EllipticGraph."""
    y, t = gr.y, gr.timestep
    labeled = y >= 0
    train = (t <= t_train_end) & labeled
    if train_window_k is not None:
        t_lo = max(1, t_train_end - int(train_window_k) + 1)
        train = (t >= t_lo) & (t <= t_train_end) & labeled
    gr.train_mask = train
    gr.val_mask = (t > t_train_end) & (t <= t_val_end) & labeled
    gr.test_mask = (t > t_val_end) & labeled
    return gr


def replicate(gr: EllipticGraph, k: int, seed: int = 42) -> EllipticGraph:
    """Block-diagonal k copies; node ids offset by r*N, features re-drawn per replica
    (seed+r), timestep values kept 1..T so the sharding unit is (replica, timestep)."""
    n = gr.num_nodes
    xs, eis = [gr.x], [gr.edge_index]
    for r in range(1, k):
        g = torch.Generator().manual_seed(seed + r)
        xs.append(_features(n, gr.x.size(1), g))
        eis.append(gr.edge_index + r * n)
    rep = lambda t: None if t is None else t.repeat(k)
    return EllipticGraph(x=torch.cat(xs), edge_index=torch.cat(eis, dim=1).contiguous(),
                         y=gr.y.repeat(k), timestep=gr.timestep.repeat(k),
                         train_mask=rep(gr.train_mask), val_mask=rep(gr.val_mask),
                         test_mask=rep(gr.test_mask))


def adversarial_tiny() -> EllipticGraph:
    """Hand-built edge cases: duplicate (0,1), reciprocal (0,1)/(1,0),
    self-loop (2,2) twice, an isolated node (7), and a hub (node 8) with in-degree 500."""
    base = torch.tensor([[0, 0, 1, 2, 2, 3, 4, 5, 6], [1, 1, 0, 2, 2, 4, 3, 6, 5]])
    hub_src = torch.arange(9, 509)
    hub = torch.stack([hub_src, torch.full_like(hub_src, 8)])
    ei = torch.cat([base, hub], dim=1)
    n = 509
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, 12, generator=g)
    y = torch.randint(-1, 2, (n,), generator=g)
    t = torch.ones(n, dtype=torch.int64)
    out = EllipticGraph(x=x, edge_index=ei, y=y, timestep=t)
    return set_temporal_masks(out, 1, 1)
