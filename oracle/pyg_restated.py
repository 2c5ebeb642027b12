"""CPU oracle for the GNN message-passing hot path  --  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference (`/root/reference/src/models/gnn.py:8`) imports
GCNConv / SAGEConv / GATConv from `torch_geometric` (un-vendored third-party
dependency; unpinned in `environment.yml:27`, 2.5.3 in `.github/workflows/ci.yml:17`).
torch_geometric is not installable in this image, and the reference's own tests
(`tests/test_masks_and_metrics.py:8-28`) never construct a conv, so nothing the
reference ships pins conv outputs.  This file restates the *published* PyG 2.5.3
algorithm for the three convs in plain torch CPU ops -- the
same ATen ops PyG dispatches (`index_select`, `scatter_add_`, `scatter_reduce_`,
`addmm`, `batch_norm`) -- and composes them exactly as `src/models/gnn.py` does.

What pins it instead:
  * hand-derived known-answer vectors in `tests/test_oracle_known_answers.py`;
  * an independent float64 dense-adjacency evaluation of the three layer equations (values and
    gradients, 1e-12) in `tests/test_oracle_dense_algebra.py`;
  * `tests/golden/make_golden.py` imports the reference's *own* `src/models/gnn.py` and
    `src/train_gnn.py` (unmodified, from /root/reference) with these restated convs
    injected under the name `torch_geometric.nn`, and records what the reference's
    nets / `train_epoch` produce; the fixtures are committed under `tests/golden/`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module.  The product package never does.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# torch_geometric.utils.scatter (native-torch path)           SURVEY.md App. A
# ----------------------------------------------------------------------------


def _bcast(index: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    # PyG `broadcast(index, src, dim=0)`: view(-1,1,...).expand_as(src)
    size = [1] * src.dim()
    size[0] = -1
    return index.view(size).expand_as(src)


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    size = list(src.size())
    size[0] = dim_size
    return src.new_zeros(size).scatter_add_(0, _bcast(index, src), src)


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    count = src.new_zeros(dim_size)
    count.scatter_add_(0, index, src.new_ones(src.size(0)))
    count = count.clamp(min=1)
    out = scatter_sum(src, index, dim_size)
    return out / _bcast(count, out)


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    size = list(src.size())
    size[0] = dim_size
    return src.new_zeros(size).scatter_reduce_(
        0, _bcast(index, src), src, reduce="amax", include_self=False
    )


# ----------------------------------------------------------------------------
# self-loop utilities and gcn_norm                                  App. A.1
# ----------------------------------------------------------------------------


def add_remaining_self_loops(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """Non-loop edges in original order, then (0,0)...(N-1,N-1)."""
    mask = edge_index[0] != edge_index[1]
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index[:, mask], loop], dim=1)


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32):
    ei2 = add_remaining_self_loops(edge_index, num_nodes)
    w = torch.ones(ei2.size(1), dtype=dtype, device=edge_index.device)
    row, col = ei2[0], ei2[1]
    deg = torch.zeros(num_nodes, dtype=dtype, device=edge_index.device)
    deg.scatter_add_(0, col, w)
    dis = deg.pow_(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    w = dis[row] * w * dis[col]
    return ei2, w


# ----------------------------------------------------------------------------
# PyG `nn.dense.linear.Linear` initialisers
# ----------------------------------------------------------------------------


def _glorot_(t: torch.Tensor):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


def _kaiming_uniform_lin_(t: torch.Tensor, fan: int):
    # PyG Linear(weight_initializer=None): kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan))
    bound = 1.0 / math.sqrt(fan) if fan > 0 else 0.0
    with torch.no_grad():
        t.uniform_(-bound, bound)


class _PyGLinear(nn.Module):
    """`torch_geometric.nn.dense.linear.Linear`: parameters `weight` [out,in], `bias` [out]."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, init: str = "kaiming"):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self._init = init
        self.reset_parameters()

    def reset_parameters(self):
        if self._init == "glorot":
            _glorot_(self.weight)
        else:
            _kaiming_uniform_lin_(self.weight, self.in_channels)
        if self.bias is not None:
            if self._init == "glorot":
                with torch.no_grad():
                    self.bias.zero_()
            else:
                _kaiming_uniform_lin_(self.bias, self.in_channels)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


# ----------------------------------------------------------------------------
# The three convs                                          App. A.1 / A.2 / A.3
# ----------------------------------------------------------------------------


class GCNConv(nn.Module):
    """PyG GCNConv(in, out) with defaults (improved=False, cached=False,
    add_self_loops=True, normalize=True, bias=True).  Call site: gnn.py:20-23,28,31."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _PyGLinear(in_channels, out_channels, bias=False, init="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        n = x.size(0)
        ei2, w = gcn_norm(edge_index, n, dtype=x.dtype)  # recomputed on every call
        h = self.lin(x)
        msg = w.view(-1, 1) * h.index_select(0, ei2[0])
        out = scatter_sum(msg, ei2[1], n)
        return out + self.bias


class SAGEConv(nn.Module):
    """PyG SAGEConv(in, out): aggr='mean', root_weight=True, bias=True.
    Call site: gnn.py:41-44,49,52 and 125-128,187,193."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = _PyGLinear(in_channels, out_channels, bias=True)
        self.lin_r = _PyGLinear(in_channels, out_channels, bias=False)
        # PyG re-runs reset_parameters() at the end of __init__ (RNG consumed twice)
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def aggregate(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        return scatter_mean(x.index_select(0, edge_index[0]), edge_index[1], x.size(0))

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        m = self.aggregate(x, edge_index)
        return self.lin_l(m) + self.lin_r(x)


class GATConv(nn.Module):
    """PyG >= 2.5 GATConv(in, C, heads=H, concat=...), negative_slope=0.2, dropout=0,
    add_self_loops=True, bias=True.  Call site: gnn.py:64-67,72,75."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.heads, self.concat, self.negative_slope = heads, concat, negative_slope
        self.lin = _PyGLinear(in_channels, heads * out_channels, bias=False, init="glorot")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        _glorot_(self.att_src)
        _glorot_(self.att_dst)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        n, H, C = x.size(0), self.heads, self.out_channels
        xs = self.lin(x).view(n, H, C)
        a_s = (xs * self.att_src).sum(dim=-1)
        a_d = (xs * self.att_dst).sum(dim=-1)
        ei2 = add_remaining_self_loops(edge_index, n)  # remove_self_loops + add_self_loops
        src, dst = ei2[0], ei2[1]
        e = F.leaky_relu(a_s.index_select(0, src) + a_d.index_select(0, dst), self.negative_slope)
        mx = scatter_max(e.detach(), dst, n)
        p = (e - mx.index_select(0, dst)).exp()
        den = scatter_sum(p, dst, n) + 1e-16
        alpha = p / den.index_select(0, dst)
        out = scatter_sum(alpha.unsqueeze(-1) * xs.index_select(0, src), dst, n)
        out = out.view(n, H * C) if self.concat else out.mean(dim=1)
        return out + self.bias


# ----------------------------------------------------------------------------
# The four nets, composed exactly as src/models/gnn.py composes them
# ----------------------------------------------------------------------------


def _dropout(h, p, training, masks, li):
    """F.dropout, or -- for trajectory parity with the CUDA path -- an injected keep-mask
    (SURVEY.md section 7 'Dropout and end-to-end parity', mode iii)."""
    if masks is not None and training and p > 0:
        keep = masks[li].to(h.dtype)
        return h * keep * (1.0 / (1.0 - p))
    return F.dropout(h, p=p, training=training)


class _StackNet(nn.Module):
    act = staticmethod(F.relu)

    def forward(self, x, edge_index, t_idx: Optional[torch.Tensor] = None,
                dropout_masks: Optional[List[torch.Tensor]] = None):
        h = x
        for li, conv in enumerate(self.convs[:-1]):
            h = conv(h, edge_index)
            h = self.act(h)
            h = _dropout(h, self.dropout, self.training, dropout_masks, li)
        return self.convs[-1](h, edge_index)


class GCNNet(_StackNet):  # gnn.py:14-32
    def __init__(self, in_dim, hidden_dim=128, layers=3, dropout=0.2, num_classes=2):
        super().__init__()
        assert layers >= 2
        self.dropout = dropout
        self.convs = nn.ModuleList([GCNConv(in_dim, hidden_dim)])
        for _ in range(layers - 2):
            self.convs.append(GCNConv(hidden_dim, hidden_dim))
        self.convs.append(GCNConv(hidden_dim, num_classes))


class SAGENet(_StackNet):  # gnn.py:35-53
    def __init__(self, in_dim, hidden_dim=128, layers=3, dropout=0.2, num_classes=2):
        super().__init__()
        assert layers >= 2
        self.dropout = dropout
        self.convs = nn.ModuleList([SAGEConv(in_dim, hidden_dim)])
        for _ in range(layers - 2):
            self.convs.append(SAGEConv(hidden_dim, hidden_dim))
        self.convs.append(SAGEConv(hidden_dim, num_classes))


class GATNet(_StackNet):  # gnn.py:56-76
    act = staticmethod(F.elu)

    def __init__(self, in_dim, hidden_dim=128, layers=3, dropout=0.2, num_classes=2, heads=4):
        super().__init__()
        assert layers >= 2
        self.dropout = dropout
        self.convs = nn.ModuleList([GATConv(in_dim, hidden_dim // heads, heads=heads)])
        for _ in range(layers - 2):
            self.convs.append(GATConv(hidden_dim, hidden_dim // heads, heads=heads))
        self.convs.append(GATConv(hidden_dim, num_classes, heads=1, concat=False))


def sinusoid_table(max_timestep: int, dim: int) -> torch.Tensor:
    """Rows t=1..max_timestep of SAGEResBNNet._sinusoid (gnn.py:146-166) -> [max_timestep, dim]."""
    t_idx = torch.arange(1, max_timestep + 1)
    t = torch.clamp(t_idx.long() - 1, 0, max_timestep - 1).to(torch.float32)
    t = t / max(float(max_timestep - 1), 1.0)
    half = dim // 2
    freqs = torch.arange(1, half + 1, dtype=t.dtype) * (2.0 * math.pi)
    ang = t.unsqueeze(1) * freqs.unsqueeze(0)
    feat = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    if feat.size(1) < dim:
        feat = torch.cat([feat, torch.zeros(feat.size(0), dim - feat.size(1))], dim=1)
    return feat


class SAGEResBNNet(nn.Module):  # gnn.py:82-194
    def __init__(self, in_dim, hidden_dim=128, layers=3, dropout=0.2, num_classes=2, use_bn=True,
                 residual=True, time_embed_dim=0, time_embed_type="learned", max_timestep=50):
        super().__init__()
        assert layers >= 2
        self.dropout = float(dropout)
        self.use_bn, self.residual = bool(use_bn), bool(residual)
        self.time_embed_dim, self.time_embed_type = int(time_embed_dim), str(time_embed_type)
        self.max_timestep = int(max_timestep)
        self.time_emb = None
        if self.time_embed_dim > 0 and self.time_embed_type == "learned":
            self.time_emb = nn.Embedding(self.max_timestep, self.time_embed_dim)
            in_dim = in_dim + self.time_embed_dim
        elif self.time_embed_dim > 0 and self.time_embed_type == "sin":
            in_dim = in_dim + self.time_embed_dim
        else:
            self.time_embed_dim, self.time_embed_type = 0, "none"
        self.convs = nn.ModuleList([SAGEConv(in_dim, hidden_dim)])
        for _ in range(layers - 2):
            self.convs.append(SAGEConv(hidden_dim, hidden_dim))
        self.convs.append(SAGEConv(hidden_dim, num_classes))
        self.bns = nn.ModuleList()
        if self.use_bn:
            for _ in range(layers - 1):
                self.bns.append(nn.BatchNorm1d(hidden_dim))
        self.res_projs = nn.ModuleList()
        in_dims = [in_dim] + [hidden_dim] * (layers - 2)
        for d_in in in_dims:
            self.res_projs.append(nn.Identity() if d_in == hidden_dim
                                  else nn.Linear(d_in, hidden_dim, bias=False))

    def _inject_time(self, x, t_idx):
        if self.time_embed_dim <= 0 or t_idx is None:
            return x
        tidx = torch.clamp(t_idx.long() - 1, 0, self.max_timestep - 1)
        if self.time_embed_type == "learned":
            return torch.cat([x, self.time_emb(tidx)], dim=1)
        t = tidx.to(torch.float32) / max(float(self.max_timestep - 1), 1.0)
        half = self.time_embed_dim // 2
        freqs = torch.arange(1, half + 1, dtype=t.dtype) * (2.0 * math.pi)
        ang = t.unsqueeze(1) * freqs.unsqueeze(0)
        feat = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
        if feat.size(1) < self.time_embed_dim:
            feat = torch.cat([feat, torch.zeros(feat.size(0), self.time_embed_dim - feat.size(1))], 1)
        return torch.cat([x, feat], dim=1)

    def forward(self, x, edge_index, t_idx: Optional[torch.Tensor] = None,
                dropout_masks: Optional[List[torch.Tensor]] = None):
        x = self._inject_time(x, t_idx)
        h = x
        for li, conv in enumerate(self.convs[:-1]):
            h_in = h
            h = conv(h, edge_index)
            if self.use_bn:
                h = self.bns[li](h)
            h = F.relu(h)
            h = _dropout(h, self.dropout, self.training, dropout_masks, li)
            # NOTE gnn.py:192 adds the projection unconditionally (`residual` is stored, never read)
            h = h + self.res_projs[li](h_in)
        return self.convs[-1](h, edge_index)


def build_model(arch: str, in_dim: int, cfg: dict) -> nn.Module:  # train_gnn.py:67-104
    if arch == "gcn":
        return GCNNet(in_dim, hidden_dim=cfg["hidden_dim"], layers=cfg["layers"], dropout=cfg["dropout"])
    if arch == "sage":
        return SAGENet(in_dim, hidden_dim=cfg["hidden_dim"], layers=cfg["layers"], dropout=cfg["dropout"])
    if arch == "gat":
        return GATNet(in_dim, hidden_dim=cfg["hidden_dim"], layers=cfg["layers"],
                      heads=cfg.get("heads", 4), dropout=cfg["dropout"])
    if arch in ("sage_resbn", "sage_bn", "sage_res"):
        return SAGEResBNNet(in_dim, hidden_dim=cfg.get("hidden_dim", 128), layers=cfg.get("layers", 3),
                            dropout=cfg.get("dropout", 0.2), num_classes=2,
                            use_bn=cfg.get("use_bn", True), residual=cfg.get("residual", True),
                            time_embed_dim=cfg.get("time_embed_dim", 0),
                            time_embed_type=cfg.get("time_embed_type", "learned"),
                            max_timestep=cfg.get("max_timestep", 49))
    raise ValueError("Unknown arch")


# ----------------------------------------------------------------------------
# The train / eval step, restating src/train_gnn.py:116-123,136-183,187-209,248-257
# ----------------------------------------------------------------------------


def make_temporal_masks(y, t, t_train_end, t_val_end, train_window_k=None):
    """dataset_elliptic.py:268-290 on bare tensors -> (train, val, test) bool masks."""
    labeled = y >= 0
    train = (t <= t_train_end) & labeled
    val = (t > t_train_end) & (t <= t_val_end) & labeled
    test = (t > t_val_end) & labeled
    if train_window_k is not None:
        t_lo = max(1, t_train_end - train_window_k + 1)
        train = (t >= t_lo) & (t <= t_train_end) & labeled
    return train, val, test


def class_weight(train_y: torch.Tensor) -> torch.Tensor:
    pos = (train_y == 1).sum().item()
    neg = (train_y == 0).sum().item()
    if pos == 0 or neg == 0:
        return torch.tensor([1.0, 1.0], dtype=torch.float32)
    return torch.tensor([(pos + neg) / (2.0 * neg), (pos + neg) / (2.0 * pos)], dtype=torch.float32)


def masked_weighted_ce(logits, y, train_mask, cw):
    """`F.cross_entropy(..., weight=cw, reduction='none').mean()` -- an UNWEIGHTED mean of
    weighted per-sample losses (train_gnn.py:159-176)."""
    lv = F.cross_entropy(logits[train_mask], y[train_mask], weight=cw, reduction="none")
    return lv.mean()


def make_loss_fn(cfg, cw, model, t_min, t_max):
    """`_make_loss_fn` (train_gnn.py:136-183): focal / class-weighted CE, optional linear|sqrt time weighting, optional
    L2 on the learned time table.  Pinned against the reference's own function (tests/golden/make_loss_golden.py)."""
    scheme = str(cfg.get("time_loss_weighting", "none"))
    embed_l2 = float(cfg.get("time_embed_l2", 0.0))
    use_focal = bool(cfg.get("focal_loss", False))
    gamma = float(cfg.get("focal_gamma", 2.0))

    def loss_fn(logits, target, t_idx=None):
        if use_focal:
            ce = F.cross_entropy(logits, target, reduction="none")
            pt = torch.softmax(logits, dim=1)[torch.arange(len(target)), target]
            lv = ((1 - pt) ** gamma) * ce
        else:
            lv = F.cross_entropy(logits, target, weight=cw, reduction="none")
        if scheme != "none" and t_idx is not None:
            wt = (t_idx.float() - float(t_min)) / max(float(t_max - t_min), 1.0)
            if scheme == "sqrt":
                wt = torch.sqrt(torch.clamp(wt, min=0.0))
            elif scheme != "linear":
                raise ValueError(f"unknown time_loss_weighting={scheme}")
            lv = lv * torch.clamp(wt, min=1e-3)
        loss = lv.mean()
        if embed_l2 > 0.0 and getattr(model, "time_emb", None) is not None:
            loss = loss + embed_l2 * model.time_emb.weight.pow(2).mean()
        return loss

    return loss_fn


def train_step(model, x, edge_index, t_idx, y, train_mask, cw, optimizer, grad_clip=1.0,
               amp_dtype=None, dropout_masks=None, loss_fn=None, time_weighted=False):
    """One `train_epoch` body (train_gnn.py:187-209).  `amp_dtype=torch.bfloat16` runs the
    forward under CPU autocast(bf16) with no GradScaler."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    uses_t = getattr(model, "time_embed_dim", 0) > 0
    kw = {} if dropout_masks is None else {"dropout_masks": dropout_masks}

    def _loss(lg):      # train_gnn.py:196-201: t_idx only when time weighting is configured
        if loss_fn is None:
            return masked_weighted_ce(lg, y, train_mask, cw)
        return loss_fn(lg[train_mask], y[train_mask], t_idx[train_mask] if time_weighted else None)

    if amp_dtype is not None:
        with torch.autocast(device_type="cpu", dtype=amp_dtype):
            logits = model(x, edge_index, t_idx if uses_t else None, **kw)
            loss = _loss(logits.float())
    else:
        logits = model(x, edge_index, t_idx if uses_t else None, **kw)
        loss = _loss(logits)
    loss.backward()
    if grad_clip and grad_clip > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip)
    optimizer.step()
    return float(loss.item()), logits.detach()


@torch.no_grad()
def eval_probs(model, x, edge_index, t_idx):
    """eval_split (train_gnn.py:248-257): fp32, never under autocast."""
    model.eval()
    uses_t = getattr(model, "time_embed_dim", 0) > 0
    logits = model(x, edge_index, t_idx if uses_t else None)
    return torch.softmax(logits, dim=1)[:, 1], logits


def train_epoch_minibatch(model, batches, optimizer, loss_fn, cfg: dict) -> float:
    """`train_epoch_minibatch` (train_gnn.py:212-245) without AMP: per batch forward on the sampled subgraph, loss on
    the first `batch_size` rows (the seeds), backward, clip, step; returns the seed-weighted mean of the batch losses.
    `batches`: iterable of objects with `x`, `edge_index`, `y`, `timestep`, `batch_size` (CPU tensors)."""
    model.train()
    uses_t = getattr(model, "time_embed_dim", 0) > 0
    total_loss, total_examples = 0.0, 0
    for batch in batches:
        optimizer.zero_grad(set_to_none=True)
        bs = int(batch.batch_size)
        logits = model(batch.x, batch.edge_index, batch.timestep if uses_t else None)
        t_idx = batch.timestep[:bs] if cfg.get("time_loss_weighting", "none") != "none" else None    # :229-233
        loss = loss_fn(logits[:bs], batch.y[:bs], t_idx)
        loss.backward()
        if cfg.get("grad_clip", 0) and cfg["grad_clip"] > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["grad_clip"])
        optimizer.step()
        total_loss += loss.item() * bs                                                                # :242-243
        total_examples += bs
    return float(total_loss / total_examples) if total_examples else 0.0


@torch.no_grad()
def eval_val_minibatch(model, batches):
    """`eval_val_minibatch` (train_gnn.py:261-280): (labels, P(illicit)) of the seeds of every batch, concatenated."""
    model.eval()
    uses_t = getattr(model, "time_embed_dim", 0) > 0
    ys, ps = [], []
    for batch in batches:
        bs = int(batch.batch_size)
        logits = model(batch.x, batch.edge_index, batch.timestep if uses_t else None)[:bs]
        ps.append(torch.softmax(logits, dim=1)[:, 1])
        ys.append(batch.y[:bs])
    if not ys:
        return torch.empty(0, dtype=torch.long), torch.empty(0)
    return torch.cat(ys), torch.cat(ps)
