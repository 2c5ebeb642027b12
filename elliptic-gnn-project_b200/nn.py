"""Drop-in `GCNConv`, `SAGEConv`, `GATConv` with PyG's constructor signature, call signature
`conv(x, edge_index)` and parameter names (state-dict compatible with a PyG-trained
`best.ckpt`, `/root/reference/src/train_gnn.py:522`), backed by the sm_100a kernels.
Call sites replaced: `/root/reference/src/models/gnn.py:20-23,41-44,64-67,125-128`.
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch
import torch.nn as nn

from . import ops
from .graph import Graph, cached_graph


def _glorot_(t: torch.Tensor):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


class _Lin(nn.Module):
    """Parameter holder named like torch_geometric.nn.dense.linear.Linear (`weight`, `bias`)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool, glorot: bool = False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.glorot = glorot
        self.reset_parameters()

    def reset_parameters(self):
        if self.glorot:
            _glorot_(self.weight)
        else:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            with torch.no_grad():
                self.weight.uniform_(-bound, bound)
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            with torch.no_grad():
                self.bias.uniform_(-bound, bound)


def _graph_of(edge_index: Union[torch.Tensor, Graph], n: int, self_loops: bool) -> Graph:
    if isinstance(edge_index, Graph):
        return edge_index
    return cached_graph(edge_index, n, self_loops=self_loops)


def _check_input(x: torch.Tensor, in_channels: int) -> torch.Tensor:
    ops.set_f32_tc()
    if not x.is_cuda:
        raise RuntimeError("egnn_b200 convs run on CUDA tensors only (no CPU fallback)")
    if x.dim() != 2 or x.size(1) != in_channels:
        raise ValueError(f"expected x of shape [N, {in_channels}], got {tuple(x.shape)}")
    return ops.widen_fp16(x)          # fp16 from a caller's fp16-autocast modules: computed in fp32 (ops.amp_bf16)


class SAGEConv(nn.Module):
    """`SAGEConv(in_channels, out_channels)`: mean aggregation, root weight, bias."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = _Lin(in_channels, out_channels, bias=True)
        self.lin_r = _Lin(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def forward(self, x: torch.Tensor, edge_index: Union[torch.Tensor, Graph]) -> torch.Tensor:
        x = _check_input(x, self.in_channels)
        g = _graph_of(edge_index, x.size(0), self_loops=False)
        if ops.sage_out_supported(x, self.out_channels):
            # narrow output (the logits layer): project first, aggregate at width out_channels
            return ops.SageOutFn.apply(ops._rows(x), self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, g)
        return ops.SageConvFn.apply(x, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, None, g,
                                    ops.amp_bf16())

    def forward_with_res(self, x: torch.Tensor, edge_index: Union[torch.Tensor, Graph], res_weight: torch.Tensor):
        """(conv(x, edge_index), x @ res_weight.T): SAGEResBNNet's residual projection
        (`src/models/gnn.py:141-144,192`) folded into the conv's GEMM."""
        x = _check_input(x, self.in_channels)
        g = _graph_of(edge_index, x.size(0), self_loops=False)
        return ops.SageConvFn.apply(x, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, res_weight, g,
                                    ops.amp_bf16())


class GCNConv(nn.Module):
    """`GCNConv(in_channels, out_channels)`: add_self_loops, symmetric normalisation, bias."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels, bias=False, glorot=True)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        # set by egnn_b200.models.GCNNet on its hidden layers: emit bf16 under bf16 autocast (PyG's fp32 output is
        # rounded to bf16 by the next layer's Linear anyway); a standalone GCNConv keeps PyG's fp32 output
        self._hidden_bf16 = False

    def reset_parameters(self):
        self.lin.reset_parameters()
        with torch.no_grad():
            self.bias.zero_()

    def forward(self, x: torch.Tensor, edge_index: Union[torch.Tensor, Graph]) -> torch.Tensor:
        x = _check_input(x, self.in_channels)
        g = _graph_of(edge_index, x.size(0), self_loops=True)
        if self.out_channels in (2, 4) and ops.sage_out_supported(x, self.out_channels):
            # narrow output (the logits layer): project first, aggregate at width out_channels
            return ops.GcnOutFn.apply(ops._rows(x), self.lin.weight, self.bias, g)
        return ops.GcnConvFn.apply(x, self.lin.weight, self.bias, g, ops.amp_bf16(), self._hidden_bf16)


class GATConv(nn.Module):
    """`GATConv(in_channels, out_channels, heads=1, concat=True)`, negative_slope 0.2,
    attention dropout 0 (the reference never passes `dropout=`), self loops, bias."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 bias: bool = True):
        super().__init__()
        if dropout != 0.0 or not add_self_loops or not bias:
            raise NotImplementedError("only the GATConv configuration used by the reference is implemented "
                                      "(dropout=0, add_self_loops=True, bias=True)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.heads, self.concat, self.negative_slope = heads, concat, negative_slope
        self.lin = _Lin(in_channels, heads * out_channels, bias=False, glorot=True)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        _glorot_(self.att_src)
        _glorot_(self.att_dst)

    def forward(self, x: torch.Tensor, edge_index: Union[torch.Tensor, Graph]) -> torch.Tensor:
        x = _check_input(x, self.in_channels)
        g = _graph_of(edge_index, x.size(0), self_loops=True)
        return ops.GatConvFn.apply(x, self.lin.weight, self.att_src, self.att_dst, self.bias, g, self.heads,
                                   self.out_channels, self.concat, self.negative_slope, ops.amp_bf16())
