"""The four networks of `/root/reference/src/models/gnn.py` with identical constructors,
`forward(x, edge_index, t_idx=None)` signature, module/parameter names and config toggles,
re-assembled from the B200 ops: conv -> fused (BN) + activation + dropout (+ residual).
`build_model` mirrors `/root/reference/src/train_gnn.py:67-104`.
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch
import torch.nn as nn

from . import fused, ops
from ._lib import ACT_ELU, ACT_RELU
from .graph import Graph, cached_graph
from .nn import GATConv, GCNConv, SAGEConv


class _DropoutMixin:
    """Philox dropout stream shared by a net's layers (seeded from torch's RNG at first use, so
    `set_seed` of the reference, `src/utils/common.py:11-17`, still controls it)."""

    _drop: Optional[ops.DropoutState] = None
    stats_reducer: Optional[ops.StatsReducer] = None
    row0: int = 0  # global id of this shard's first node (timestep-sharded runs)

    def dropout_state(self, device) -> ops.DropoutState:
        if self._drop is None or self._drop.offset.device != device:
            seed = int(torch.randint(0, 2**62, (1,)).item())
            self._drop = ops.DropoutState(seed, device)
        return self._drop

    def set_dropout_seed(self, seed: int, device="cuda"):
        self._drop = ops.DropoutState(seed, torch.device(device))


class _StackNet(nn.Module, _DropoutMixin):
    _act = ACT_RELU

    def forward(self, x, edge_index: Union[torch.Tensor, Graph], t_idx: Optional[torch.Tensor] = None):
        x = ops.widen_fp16(x)
        if getattr(self, "no_residual", False) and type(self).forward is _StackNet.forward:
            bf16 = ops.amp_bf16()
            if fused.supported(self, x, bf16) and (bf16 or ops.set_f32_tc()):
                g = edge_index if isinstance(edge_index, Graph) else cached_graph(edge_index, x.size(0))
                return fused.SageResBNFn.apply(self, x, g, None, bf16, *fused.param_order(self))
        h = x
        drop = None
        if self.training and self.dropout > 0:
            drop = self.dropout_state(x.device)
            drop.advance()
        for li, conv in enumerate(self.convs[:-1]):
            h = conv(h, edge_index)
            h = ops.act_dropout(h, self._act, self.dropout, self.training, drop, li, self.row0)
        return self.convs[-1](h, edge_index)


class GCNNet(_StackNet):
    def __init__(self, in_dim, hidden_dim=128, layers=3, dropout=0.2, num_classes=2):
        super().__init__()
        assert layers >= 2
        self.dropout = dropout
        self.convs = nn.ModuleList([GCNConv(in_dim, hidden_dim)])
        for _ in range(layers - 2):
            self.convs.append(GCNConv(hidden_dim, hidden_dim))
        self.convs.append(GCNConv(hidden_dim, num_classes))
        for conv in self.convs[:-1]:
            conv._hidden_bf16 = True   # hidden activations in bf16 under bf16 autocast (see nn.GCNConv)


class SAGENet(_StackNet):
    """`SAGENet` (`/root/reference/src/models/gnn.py:35-53`): h = dropout(relu(SAGEConv(h))) per hidden layer.  Shapes the
    explicit kernel sequence serves (fused.py: hidden % 16 == 0, <= 128) run through it -- the same sequence as
    SAGE-ResBN with BatchNorm, residual and time features switched off; everything else takes the per-op path."""
    # what fused.py reads from a net, fixed for this architecture
    use_bn, no_residual, time_embed_dim, time_embed_type, time_emb, max_timestep = False, True, 0, "none", None, 0
    bns, res_projs = (), ()

    def __init__(self, in_dim, hidden_dim=128, layers=3, dropout=0.2, num_classes=2):
        super().__init__()
        assert layers >= 2
        self.dropout = dropout
        self.in_dim = in_dim
        self.convs = nn.ModuleList([SAGEConv(in_dim, hidden_dim)])
        for _ in range(layers - 2):
            self.convs.append(SAGEConv(hidden_dim, hidden_dim))
        self.convs.append(SAGEConv(hidden_dim, num_classes))


class GATNet(_StackNet):
    _act = ACT_ELU

    def __init__(self, in_dim, hidden_dim=128, layers=3, dropout=0.2, num_classes=2, heads=4):
        super().__init__()
        assert layers >= 2
        self.dropout = dropout
        self.convs = nn.ModuleList([GATConv(in_dim, hidden_dim // heads, heads=heads)])
        for _ in range(layers - 2):
            self.convs.append(GATConv(hidden_dim, hidden_dim // heads, heads=heads))
        self.convs.append(GATConv(hidden_dim, num_classes, heads=1, concat=False))


class _ResProj(nn.Linear):
    """`nn.Linear(d_in, d_out, bias=False)` residual projection running on the egnn GEMM."""

    def forward(self, x):
        bf16 = ops.amp_bf16()
        return _LinearFn.apply(ops.widen_fp16(x), self.weight, bf16)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bf16: bool):
        cd = torch.bfloat16 if bf16 else torch.float32
        xg = ops.to_compute(x, cd)
        wc = w if w.dtype == cd else ops.cast(w, cd)
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(xg, wc)
        return ops.linear_fwd(xg, wc, out_dtype=cd)

    @staticmethod
    def backward(ctx, dy):
        xg, wc = ctx.saved_tensors
        dy = ops._rows(dy)
        if dy.dtype != xg.dtype:
            dy = ops.cast(dy, xg.dtype)
        dw = ops.linear_wgrad(dy, xg)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear_dgrad(dy, wc)
            if dx.dtype != ctx.x_dtype:
                dx = ops.cast(dx, ctx.x_dtype)
        return dx, dw, None


def sinusoid_table(max_timestep: int, dim: int) -> torch.Tensor:
    """The 49 distinct rows `SAGEResBNNet._sinusoid` (gnn.py:146-166) can produce, computed on the
    host with the reference's own torch ops so the device LUT is bit-identical to the CPU path."""
    t = torch.arange(0, max_timestep, dtype=torch.float32)  # == clamp(t_idx - 1, 0, T-1)
    t = t / max(float(max_timestep - 1), 1.0)
    half = dim // 2
    freqs = torch.arange(1, half + 1, dtype=t.dtype) * (2.0 * math.pi)
    angles = t.unsqueeze(1) * freqs.unsqueeze(0)
    feat = torch.cat([torch.sin(angles), torch.cos(angles)], dim=1)
    if feat.size(1) < dim:
        feat = torch.cat([feat, torch.zeros((feat.size(0), dim - feat.size(1)))], dim=1)
    return feat


class SAGEResBNNet(nn.Module, _DropoutMixin):
    """SAGE with residual connections, BatchNorm and an optional timestep embedding
    (`/root/reference/src/models/gnn.py:82-194`)."""

    def __init__(self, in_dim, hidden_dim=128, layers=3, dropout=0.2, num_classes=2, use_bn=True,
                 residual=True, time_embed_dim=0, time_embed_type="learned", max_timestep=50):
        super().__init__()
        assert layers >= 2
        self.dropout = float(dropout)
        self.use_bn, self.residual = bool(use_bn), bool(residual)
        self.time_embed_dim, self.time_embed_type = int(time_embed_dim), str(time_embed_type)
        self.max_timestep = int(max_timestep)
        self.raw_in_dim = in_dim
        self.time_emb = None
        if self.time_embed_dim > 0 and self.time_embed_type == "learned":
            self.time_emb = nn.Embedding(self.max_timestep, self.time_embed_dim)
            in_dim = in_dim + self.time_embed_dim
        elif self.time_embed_dim > 0 and self.time_embed_type == "sin":
            in_dim = in_dim + self.time_embed_dim
            self.register_buffer("_sin_table", sinusoid_table(self.max_timestep, self.time_embed_dim),
                                 persistent=False)
        else:
            self.time_embed_dim, self.time_embed_type = 0, "none"
        self.in_dim = in_dim
        self.convs = nn.ModuleList([SAGEConv(in_dim, hidden_dim)])
        for _ in range(layers - 2):
            self.convs.append(SAGEConv(hidden_dim, hidden_dim))
        self.convs.append(SAGEConv(hidden_dim, num_classes))
        self.bns = nn.ModuleList()
        if self.use_bn:
            for _ in range(layers - 1):
                self.bns.append(nn.BatchNorm1d(hidden_dim))
        self.res_projs = nn.ModuleList()
        for d_in in [in_dim] + [hidden_dim] * (layers - 2):
            self.res_projs.append(nn.Identity() if d_in == hidden_dim else _ResProj(d_in, hidden_dim, bias=False))

    def _inject_time(self, x, t_idx):
        if self.time_embed_dim <= 0 or t_idx is None:
            return x
        table = self.time_emb.weight if self.time_embed_type == "learned" else self._sin_table
        return ops.inject_time(x, t_idx, table, self.in_dim)

    def forward(self, x, edge_index: Union[torch.Tensor, Graph], t_idx: Optional[torch.Tensor] = None):
        x = ops.widen_fp16(x)
        bf16 = ops.amp_bf16()
        if fused.supported(self, x, bf16) and (bf16 or ops.set_f32_tc()):
            # the whole net as one explicit kernel sequence (fused.py): bf16 operands under autocast, 3xTF32 otherwise
            g = edge_index if isinstance(edge_index, Graph) else cached_graph(edge_index, x.size(0))
            return fused.SageResBNFn.apply(self, x, g, t_idx, bf16, *fused.param_order(self))
        x = self._inject_time(x, t_idx)
        h = x
        drop = None
        if self.training and self.dropout > 0:
            drop = self.dropout_state(x.device)
            drop.advance()
        for li, conv in enumerate(self.convs[:-1]):
            h_in = h
            proj = self.res_projs[li]
            if isinstance(proj, nn.Identity):
                z, res = conv(h, edge_index), h_in
            else:   # residual projection shares the conv's GEMM (one pass over h_in)
                z, res = conv.forward_with_res(h, edge_index, proj.weight)
            if self.use_bn:
                bn = self.bns[li]
                if self.training and bn.track_running_stats:
                    bn.num_batches_tracked += 1
                h = ops.BnActDropResFn.apply(z, res, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                             self.training, ACT_RELU, self.dropout, drop, li, self.row0,
                                             bn.eps, bn.momentum, self.stats_reducer)
            else:
                h = ops.BnActDropResFn.apply(z, res, None, None, None, None, self.training, ACT_RELU,
                                             self.dropout, drop, li, self.row0, 0.0, 0.0, None)
        return self.convs[-1](h, edge_index)


def build_model(arch: str, in_dim: int, cfg: dict) -> nn.Module:
    if arch == "gcn":
        return GCNNet(in_dim, hidden_dim=cfg["hidden_dim"], layers=cfg["layers"], dropout=cfg["dropout"])
    elif arch == "sage":
        return SAGENet(in_dim, hidden_dim=cfg["hidden_dim"], layers=cfg["layers"], dropout=cfg["dropout"])
    elif arch == "gat":
        return GATNet(in_dim, hidden_dim=cfg["hidden_dim"], layers=cfg["layers"], heads=cfg.get("heads", 4),
                      dropout=cfg["dropout"])
    elif arch in ("sage_resbn", "sage_bn", "sage_res"):
        return SAGEResBNNet(in_dim, hidden_dim=cfg.get("hidden_dim", 128), layers=cfg.get("layers", 3),
                            dropout=cfg.get("dropout", 0.2), num_classes=2, use_bn=cfg.get("use_bn", True),
                            residual=cfg.get("residual", True), time_embed_dim=cfg.get("time_embed_dim", 0),
                            time_embed_type=cfg.get("time_embed_type", "learned"),
                            max_timestep=cfg.get("max_timestep", 49))
    else:
        raise ValueError("Unknown arch")
