"""The caller side of the hot path: one full-batch `train_epoch` body and `eval_split`
(`/root/reference/src/train_gnn.py:187-209,248-257`), `class_weight` (`:116-123`), the
masked weighted CE (`:136-183`), global-norm clipping + Adam (`:203-207,357-359`), and a
CUDA-graph capture of the whole step (the reference's step is ~50 tiny launches plus a
`loss.item()` sync; on B200 the step is launch-bound unless replayed as one graph).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import check, lib, ptr, stream


def class_weight(train_y: torch.Tensor) -> torch.Tensor:
    pos = (train_y == 1).sum().item()
    neg = (train_y == 0).sum().item()
    if pos == 0 or neg == 0:
        return torch.tensor([1.0, 1.0], dtype=torch.float32)
    w_pos = (pos + neg) / (2.0 * pos)
    w_neg = (pos + neg) / (2.0 * neg)
    return torch.tensor([w_neg, w_pos], dtype=torch.float32)


def model_uses_time_embed(model) -> bool:
    return getattr(model, "time_embed_dim", 0) > 0


class FlatClipAdam:
    """clip_grad_norm_(params, max_norm) + torch.optim.Adam (coupled L2 weight decay) as ONE
    pass over a flat fp32 buffer: parameters and their .grad become views of two flat
    tensors, so the multi-GPU gradient all-reduce is also a single call on `flat_grad`."""

    def __init__(self, params, lr: float, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_norm: float = 0.0):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.ws = torch.empty(lib().egnn_adam_workspace_floats(n), dtype=torch.float32, device=dev)
        off = 0
        self.views = []
        for p in self.params:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p.data)
            self.views.append(self.flat_grad[off:off + k].view_as(p.data))
            p.grad = self.views[-1]
            off += k
        self.lr, self.wd, self.betas, self.eps, self.max_norm = lr, weight_decay, betas, eps, max_norm

    def zero_grad(self, set_to_none: bool = True):
        """Drop the .grad references: autograd then hands each parameter its gradient tensor as is
        (no `grad += new` kernel per parameter); `gather_grads()` packs them into the flat buffer."""
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        """flat_grad <- the gradients autograd produced, in ONE multi-tensor copy; afterwards every
        p.grad is a view of the flat buffer again (what clip / Adam / the all-reduce operate on)."""
        src, dst, missing = [], [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                missing.append(v)
            else:
                src.append(p.grad.reshape(v.shape).to(torch.float32))
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        for v in missing:
            v.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def step(self):
        check(lib().egnn_clip_adam_step(ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg),
                                        ptr(self.exp_avg_sq), self.n, self.lr, self.betas[0], self.betas[1],
                                        self.eps, self.wd, self.max_norm, ptr(self.step_count),
                                        ptr(self.grad_norm), ptr(self.ws), stream()))


class TrainStep:
    """One reference `train_epoch` body on resident device tensors.

    forward over ALL nodes -> masked weighted CE on the train rows -> backward -> global-norm
    clip -> Adam.  `amp=True` runs the forward under autocast(bf16) (no GradScaler: bf16 needs
    none).  `capture()` records the step into a CUDA graph; `run()` replays it.
    """

    def __init__(self, model: nn.Module, x, edge_index, timestep, y, train_mask, *, lr: float,
                 weight_decay: float, grad_clip: float = 1.0, amp: bool = False,
                 cw: Optional[torch.Tensor] = None, n_train_total: Optional[int] = None,
                 grad_reducer=None):
        self.model, self.amp = model, amp
        self.x, self.edge_index, self.timestep, self.y = x, edge_index, timestep, y
        self.train_idx = torch.nonzero(train_mask, as_tuple=False).view(-1).contiguous()  # once per run
        self.cw = (cw if cw is not None else class_weight(y[train_mask])).to(x.device)
        self.n_train_total = float(n_train_total if n_train_total is not None else self.train_idx.numel())
        self.opt = FlatClipAdam(model.parameters(), lr=lr, weight_decay=weight_decay, max_norm=grad_clip)
        self.grad_reducer = grad_reducer
        self.loss = torch.zeros((), dtype=torch.float32, device=x.device)
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    def _body(self):
        m = self.model
        m.train()
        self.opt.zero_grad()
        t = self.timestep if model_uses_time_embed(m) else None
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=self.amp):
            logits = m(self.x, self.edge_index, t)
        loss = ops.masked_weighted_ce(logits, self.y, self.train_idx, self.cw, self.n_train_total)
        loss.backward()
        self.opt.gather_grads()
        if self.grad_reducer is not None:
            self.grad_reducer(self.opt.flat_grad)
        self.opt.step()
        self.loss.copy_(loss.detach())

    def run(self) -> torch.Tensor:
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.loss

    def capture(self, warmup: int = 3):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body()
        self.graph = g
        return self


@torch.no_grad()
def eval_probs(model, x, edge_index, timestep):
    """`eval_split` forward (fp32, never under autocast): softmax(logits)[:, 1], logits."""
    model.eval()
    logits = model(x, edge_index, timestep if model_uses_time_embed(model) else None)
    return torch.softmax(logits.float(), dim=1)[:, 1], logits


def train_epoch(model, data, edge_index, optimizer, cw, cfg: dict, use_amp: bool = False) -> float:
    """Eager mirror of the reference's `train_epoch` for any torch optimizer (reference-compatible
    calling convention: `data` has x / y / timestep / train_mask on the GPU)."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=use_amp):
        logits = model(data.x, edge_index, data.timestep if model_uses_time_embed(model) else None)
    idx = torch.nonzero(data.train_mask, as_tuple=False).view(-1)
    loss = ops.masked_weighted_ce(logits, data.y, idx, cw.to(logits.device))
    loss.backward()
    if cfg.get("grad_clip", 0) and cfg["grad_clip"] > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["grad_clip"])
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return float(loss.item())
