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

from . import fused, ops
from ._lib import check, dt, lib, ptr, stream


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
                 grad_reducer=None, health_check=None, static_inputs: bool = True, loss_fn=None):
        """`train_mask` and `y` are read ONCE here (train-row indices, class weights, loss normaliser are static per
        run in the reference too: `src/train_gnn.py:301-312,362-363`); a caller that changes them builds a new
        TrainStep.  `health_check`: callable run at host synchronisation points (`loss_value()`), e.g.
        `ShardedContext.check` (raises when a peer-memory all-reduce timed out).  `static_inputs` (default): `x`,
        `timestep` and `edge_index` are not rewritten in place between steps -- the reference moves the graph to the
        device once (`src/train_gnn.py:350`) -- so layouts derived from them (sorted graph views, the layer-0 input
        `[x | time features]`) are memoised per tensor version and a captured CUDA graph may bake them in.
        `capture_dynamic()` / `run(dynamic=True)` is the variant that re-derives them every step (`HostFeed`).
        `loss_fn`: an `ops.make_loss_fn(cfg, cw, model, t_min, t_max)` callable (focal / time-weighted / embed-L2
        variants of `_make_loss_fn`, `src/train_gnn.py:136-183`); default = the plain class-weighted CE."""
        self.model, self.amp = model, amp
        self.loss_fn = loss_fn
        self.loss_spec = getattr(loss_fn, "spec", None) or dict(focal_gamma=-1.0, time_scheme=0, t_min=0.0, t_max=1.0,
                                                                embed_l2=0.0)
        self.static_inputs = bool(static_inputs)
        self.graph_dynamic: Optional[torch.cuda.CUDAGraph] = None
        self._keepalive = []
        self.health_check = health_check
        self.x, self.edge_index, self.timestep, self.y = x, edge_index, timestep, y
        self.train_mask = train_mask
        self.train_idx = torch.nonzero(train_mask, as_tuple=False).view(-1).contiguous()  # once per run
        if cw is None and loss_fn is not None:
            cw = getattr(loss_fn, "cw", None)
        self.cw = (cw if cw is not None else class_weight(y[train_mask])).to(x.device)
        self.n_train_total = float(n_train_total if n_train_total is not None else self.train_idx.numel())
        self.opt = FlatClipAdam(model.parameters(), lr=lr, weight_decay=weight_decay, max_norm=grad_clip)
        self.grad_reducer = grad_reducer
        self.loss = torch.zeros(1, dtype=torch.float32, device=x.device).squeeze(0)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._fused_views = None

    def _fused_ok(self) -> bool:
        from .models import SAGENet, SAGEResBNNet, _StackNet
        m = self.model
        return (((isinstance(m, SAGEResBNNet) and type(m).forward is SAGEResBNNet.forward)
                 or (isinstance(m, SAGENet) and type(m).forward is _StackNet.forward))
                and (self.amp or ops.F32_TC_TRAIN)      # fp32: tensor cores with exact accumulation unless opted out (ops.py)
                and fused.supported(m, self.x, self.amp) and all(p.requires_grad for p in m.parameters())
                and len(self.opt.params) == len(fused.param_order(m)))

    def _body_fused(self):
        """SAGE-ResBN, bf16: forward, loss, backward as one explicit kernel sequence (fused.py) -- no autograd graph,
        gradients written straight into the flat buffer clip + Adam read."""
        from .graph import Graph, cached_graph
        m = self.model
        m.train()
        L = lib()
        if not self.amp:
            ops.set_f32_tc(True)      # fp32 operands: exact accumulation (a gradient is taken)
        g = self.edge_index if isinstance(self.edge_index, Graph) else cached_graph(self.edge_index, self.x.size(0))
        t = self.timestep if model_uses_time_embed(m) else None
        logits, sv = fused.forward(m, self.x, g, t, True, True, self.amp)
        n = logits.size(0)
        dlog = torch.empty_like(logits)
        ws = torch.empty(L.egnn_ce_workspace_floats(self.train_idx.numel()), dtype=torch.float32, device=logits.device)
        sp = self.loss_spec
        check(L.egnn_masked_loss(ptr(logits), dt(logits), n, ptr(self.y), ptr(self.train_idx), self.train_idx.numel(),
                                 ptr(self.cw), float(self.n_train_total), sp["focal_gamma"],
                                 ptr(self.timestep) if sp["time_scheme"] else None, sp["t_min"], sp["t_max"],
                                 sp["time_scheme"], ptr(self.loss), ptr(dlog), ptr(ws), stream()))
        if self._fused_views is None:
            by_id = {id(p): v for p, v in zip(self.opt.params, self.opt.views)}
            self._fused_views = [by_id[id(p)] for p in fused.param_order(m)]
            for p, v in zip(self.opt.params, self.opt.views):
                p.grad = v
        fused.backward(m, sv, dlog, out=self._fused_views)
        if sp["embed_l2"] > 0.0 and getattr(m, "time_emb", None) is not None:   # + lambda * mean(W_time^2) (:178-180)
            w = m.time_emb.weight
            check(L.egnn_l2_mean_penalty(ptr(w), w.numel(), sp["embed_l2"], ptr(self.loss), ptr(w.grad), stream()))

    def _body(self, dynamic: bool = False):
        m = self.model
        with fused.static_inputs(self.static_inputs and not dynamic):
            if self._fused_ok():
                self._body_fused()
            else:
                m.train()
                self.opt.zero_grad()
                t = self.timestep if model_uses_time_embed(m) else None
                with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=self.amp):
                    logits = m(self.x, self.edge_index, t)
                if self.loss_fn is not None:
                    loss = self.loss_fn.on_rows(logits, self.y, self.train_idx,
                                                self.timestep if self.loss_spec["time_scheme"] else None,
                                                self.n_train_total)
                else:
                    loss = ops.masked_weighted_ce(logits, self.y, self.train_idx, self.cw, self.n_train_total)
                loss.backward()
                self.opt.gather_grads()
                self.loss.copy_(loss.detach())
        if self.grad_reducer is not None:
            self.grad_reducer(self.opt.flat_grad)
        self.opt.step()

    def run(self, dynamic: bool = False) -> torch.Tensor:
        """One step.  `dynamic=True`: the variant that assumes `x` / `timestep` were rewritten in place since the last
        step (no memoised input layouts; its own CUDA graph after `capture_dynamic()`)."""
        graph = self.graph_dynamic if dynamic else self.graph
        if graph is not None:
            graph.replay()
        else:
            self._body(dynamic)
        return self.loss

    def loss_value(self) -> float:
        """Loss of the last step on the host (one synchronisation) after the health check."""
        v = float(self.loss.item())
        if self.health_check is not None:
            self.health_check()
        return v

    def _mutable_state(self):
        """Everything a step mutates: parameters, Adam moments and step count, BatchNorm buffers, dropout offset."""
        o = self.opt
        ts = [o.flat_param, o.exp_avg, o.exp_avg_sq, o.step_count, o.grad_norm, self.loss]
        ts += [b for b in self.model.buffers() if b.is_cuda]
        drop = getattr(self.model, "_drop", None)
        if drop is not None:
            ts.append(drop.offset)
        return ts

    def capture_dynamic(self, warmup: int = 2):
        """CUDA graph of the dynamic-input variant (`run(dynamic=True)`); the optimizer / BatchNorm / dropout state is
        left as it was."""
        return self.capture(warmup=warmup, preserve_state=True, dynamic=True)

    def capture(self, warmup: int = 3, preserve_state: bool = False, dynamic: bool = False):
        """Record the step into a CUDA graph after `warmup` eager steps.  The warm-up steps are real optimizer
        steps; `preserve_state=True` puts parameters, Adam state, BatchNorm buffers and the dropout offset back to
        their values before the warm-up, so that the first replay IS step 1 of the run (the reference trains exactly
        `max_epochs` steps from the initial weights, `src/train_gnn.py:380-413`)."""
        if preserve_state and getattr(self.model, "dropout", 0) > 0 and hasattr(self.model, "dropout_state"):
            self.model.dropout_state(self.x.device)   # materialise the stream so its offset is part of the snapshot
        saved = [t.clone() for t in self._mutable_state()] if preserve_state else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body(dynamic)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body(dynamic)
        if dynamic:
            self.graph_dynamic = g
        else:
            self.graph = g
        # tensors the graph reads but did not allocate (memoised input layouts, sorted graph views) must outlive it
        from .graph import _GLOBAL_CACHE
        self._keepalive.append((list(fused.STATIC_INPUTS._d.values()), list(_GLOBAL_CACHE._d.values())))
        if saved is not None:
            for t, v in zip(self._mutable_state(), saved):
                t.copy_(v)
        return self


class HostFeed:
    """Host-resident inputs -> the device-resident step, one full-batch step per call, double-buffered.

    The reference moves the whole graph to the device once (`data.to(device)`, `src/train_gnn.py:300-312`); a
    caller whose features live on the host (streaming snapshots, a graph larger than the residency budget of a
    shared GPU) pays the PCIe copy on every step.  `submit()` enqueues the copy of the NEXT step's pinned host
    tensors into a staging set on a copy stream; `run()` moves the staged tensors into the buffers the (captured)
    step reads (device-to-device, ~40 us for the 135 MB feature matrix), rebuilds the CSR/CSC views and the row
    partition from the new `edge_index` on a side stream, runs the step and starts the loss read-back.  The copy of
    step i+1 therefore overlaps the compute of step i, and the loss of step i is read while step i+1 runs:
    steady-state time per step = max(PCIe copy, rebuild + step).  Call order: submit(); run(); submit(); run(); ... or,
    to keep the link busy across the host-side gaps, one submission ahead: submit(); submit(); run(); submit(); run(); ...
    """

    def __init__(self, step: TrainStep, host: dict, device_bufs: dict, num_nodes: int, graph):
        from .graph import build_graph, register_graph
        self._build, self._register = build_graph, register_graph
        self.step, self.host, self.dst, self.n, self.graph = step, host, device_bufs, int(num_nodes), graph
        if step.graph is not None and step.graph_dynamic is None:
            step.capture_dynamic()      # the inputs change under the step: its graph must re-derive their layouts
        for k, v in host.items():
            if not v.is_pinned():
                raise ValueError(f"host tensor '{k}' must be pinned")
        # two staging sets: the copy of submission j+1 is queued behind the copy of submission j on the copy stream
        # without waiting for run(j), so the PCIe link never idles between steps (one set left a ~0.18 ms gap per step:
        # host wake-up after the copy + the Python time of run() and submit())
        self.stage = [{k: torch.empty_like(device_bufs[k]) for k in host} for _ in range(2)]
        self.copy_stream, self.side = torch.cuda.Stream(), torch.cuda.Stream()
        self.ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_ei, self.ev_graph, self.ev_cmp = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.ev_loss = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0
        self.n_sub = self.n_run = 0
        dev = next(iter(device_bufs.values())).device
        self.ei_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ei_flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.rebuilds = 0
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
        # `TrainStep` reads the train mask and the labels ONCE (train-row indices, class weights, loss normaliser;
        # static per run in the reference too): a feed that changes them must not be served from the stale copies.
        # Their staged bytes are compared with the device buffers on every submit; a difference raises in run().
        frozen_ptrs = {t.data_ptr() for t in (step.y, getattr(step, "train_mask", None)) if t is not None}
        self.frozen = [k for k in host if device_bufs[k].data_ptr() in frozen_ptrs
                       and device_bufs[k].data_ptr() % 16 == 0]
        nf = max(len(self.frozen), 1)
        self.frozen_flag = torch.zeros((2, nf), dtype=torch.int32, device=dev)
        self.frozen_flag_host = torch.zeros((2, nf), dtype=torch.int32).pin_memory()

    def submit(self) -> None:
        """Start copying the host tensors (their current contents) for a later `run()`; at most two submissions may be
        ahead of the runs."""
        if self.n_sub - self.n_run >= 2:
            raise RuntimeError("HostFeed: two submissions are already waiting for their run()")
        slot = self.n_sub & 1
        cs = self.copy_stream
        if self.n_sub >= 2:
            cs.wait_event(self.ev_free[slot])      # the run() of two submissions ago has emptied this staging set
        elif self.n_sub == 0:
            cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            for k in sorted(self.host, key=lambda k: k != "ei"):   # edge_index first
                self.stage[slot][k].copy_(self.host[k], non_blocking=True)
            for j, k in enumerate(self.frozen):    # dst[k] never changes (a change raises), so this compare cannot race
                a, b = self.stage[slot][k], self.dst[k]
                check(lib().egnn_buffers_differ(ptr(a), ptr(b), a.numel() * a.element_size(),
                                                self.frozen_flag[slot, j:].data_ptr(), cs.cuda_stream))
            if self.frozen:
                self.frozen_flag_host[slot].copy_(self.frozen_flag[slot], non_blocking=True)
            self.ev_ready[slot].record(cs)
        self.n_sub += 1

    def run(self):
        """One step on the oldest submitted inputs; returns the loss of the PREVIOUS step (None on the first call)."""
        if self.n_run >= self.n_sub:
            raise RuntimeError("HostFeed.run() without a submit()")
        slot = self.n_run & 1
        stage = self.stage[slot]
        main = torch.cuda.current_stream()
        main.wait_event(self.ev_ready[slot])
        ei_changed = False
        if "ei" in self.dst or self.frozen:
            if "ei" in self.dst:
                # did the edge list change?  Compared HERE, on the compute stream, against the edge list the views were
                # built from as of this step (a compare at submit time could see dst["ei"] before the previous run()
                # has updated it)
                a, b = stage["ei"], self.dst["ei"]
                check(lib().egnn_buffers_differ(ptr(a), ptr(b), a.numel() * a.element_size(), ptr(self.ei_flag),
                                                main.cuda_stream))
                self.ei_flag_host.copy_(self.ei_flag, non_blocking=True)
            self.ev_cmp.record(main)
            self.ev_cmp.synchronize()          # this step's inputs have landed (their copy overlapped the last step)
            ei_changed = "ei" in self.dst and bool(int(self.ei_flag_host[0]))
            changed = [k for j, k in enumerate(self.frozen) if int(self.frozen_flag_host[slot, j])]
            if changed:
                raise RuntimeError(f"HostFeed: host tensor(s) {changed} changed, but TrainStep read the train mask / "
                                   "labels once at construction (train-row indices, class weights, loss normaliser); "
                                   "build a new TrainStep for new masks or labels")
        if ei_changed:
            self.rebuilds += 1
            self.dst["ei"].copy_(stage["ei"], non_blocking=True)
            self.ev_ei.record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.ev_ei)
                self._build(self.dst["ei"], self.n, validate=False, out=self.graph)
                self.ev_graph.record(self.side)
            self._register(self.dst["ei"], self.n, self.graph)   # eager steps look the graph up by tensor version
        for k in self.host:
            if k != "ei":
                self.dst[k].copy_(stage[k], non_blocking=True)
        self.ev_free[slot].record(main)
        self.n_run += 1
        if ei_changed:
            main.wait_event(self.ev_graph)
        self.step.run(dynamic=True)
        lslot = self.i & 1
        self.loss_host[lslot].copy_(self.step.loss, non_blocking=True)
        self.ev_loss[lslot].record(main)
        prev = None
        if self.i > 0:
            self.ev_loss[lslot ^ 1].synchronize()
            prev = float(self.loss_host[lslot ^ 1])
        self.i += 1
        return prev

    def drain(self) -> float:
        """Loss of the last step (waits for it); the next `run()` starts a new sequence."""
        slot = (self.i - 1) & 1
        self.ev_loss[slot].synchronize()
        self.i = 0
        return float(self.loss_host[slot])


@torch.no_grad()
def eval_probs(model, x, edge_index, timestep):
    """`eval_split` forward (fp32, never under autocast): softmax(logits)[:, 1], logits."""
    model.eval()
    logits = model(x, edge_index, timestep if model_uses_time_embed(model) else None)
    return torch.softmax(logits.float(), dim=1)[:, 1], logits


class EvalStep:
    """The `eval_split` forward (`src/train_gnn.py:248-257`: fp32, never under autocast, BatchNorm on its running
    statistics) over resident inputs, replayable as ONE CUDA graph: the reference runs it after every training step, so
    an eager forward -- 18 launches of 5-140 us -- is bound by the host's launch rate, not by the GPU.  `run()` returns
    the same (probs, logits) tensors on every call; parameters and BatchNorm buffers are read at replay time."""

    def __init__(self, model: nn.Module, x, edge_index, timestep):
        self.model, self.x, self.edge_index, self.timestep = model, x, edge_index, timestep
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.out = None
        self._keepalive = []

    def capture(self, warmup: int = 2):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                eval_probs(self.model, self.x, self.edge_index, self.timestep)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.out = eval_probs(self.model, self.x, self.edge_index, self.timestep)
        self.graph = g
        from .graph import _GLOBAL_CACHE
        self._keepalive.append((list(fused.STATIC_INPUTS._d.values()), list(_GLOBAL_CACHE._d.values())))
        return self

    def run(self):
        if self.graph is None:
            return eval_probs(self.model, self.x, self.edge_index, self.timestep)
        self.model.eval()          # the mode a caller observes after an eval_split call
        self.graph.replay()
        return self.out


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


def train_epoch_minibatch(model, loader, optimizer, loss_fn, cfg: dict, use_amp: bool = False) -> float:
    """Mirror of the reference's `train_epoch_minibatch` (`src/train_gnn.py:212-245`) over an `egnn_b200.NeighborLoader`
    (batches sampled and sliced on the device): per batch forward on the sampled subgraph, loss on the first
    `batch.batch_size` rows (the seeds), backward, clip, optimizer step; returns the example-weighted mean loss.
    `loss_fn(logits, target, t_idx)` as `make_loss_fn` returns it.  bf16 autocast instead of fp16 + GradScaler."""
    model.train()
    total_loss, total_examples = None, 0      # the running sum stays on the device: one read-back per epoch, not per batch
    weighting = cfg.get("time_loss_weighting", "none") != "none"
    for batch in loader:
        batch = batch.to("cuda")
        optimizer.zero_grad(set_to_none=True)
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=use_amp):
            logits = model(batch.x, batch.edge_index, batch.timestep if model_uses_time_embed(model) else None)
        bs = batch.batch_size
        t_idx = batch.timestep[:bs] if weighting else None
        loss = loss_fn(logits[:bs], batch.y[:bs], t_idx)
        loss.backward()
        if cfg.get("grad_clip", 0) and cfg["grad_clip"] > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["grad_clip"])
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
        part = loss.detach().float() * float(bs)
        total_loss = part if total_loss is None else total_loss + part
        total_examples += int(bs)
    return float(total_loss.item() / total_examples) if total_examples else 0.0


@torch.no_grad()
def eval_val_minibatch(model, loader, device=None, as_numpy: bool = True):
    """Mirror of the reference's `eval_val_minibatch` (`src/train_gnn.py:258-277`): eval-mode forward on every sampled
    subgraph of `loader`, softmax probability of the illicit class for the seeds.  Returns `(y, probs)` -- numpy arrays
    like the reference (`as_numpy=True`: one device-to-host copy at the end, not one per batch), or device tensors for
    `metrics.average_precision`.  Empty loader -> two empty arrays."""
    model.eval()
    ys, ps = [], []
    for batch in loader:
        batch = batch.to("cuda")
        logits = model(batch.x, batch.edge_index, batch.timestep if model_uses_time_embed(model) else None)
        bs = batch.batch_size
        ps.append(torch.softmax(logits[:bs].float(), dim=1)[:, 1])
        ys.append(batch.y[:bs])
    if not ys:
        import numpy as np
        return (np.array([]), np.array([])) if as_numpy else (torch.empty(0), torch.empty(0))
    y, p = torch.cat(ys), torch.cat(ps)
    return (y.cpu().numpy(), p.cpu().numpy()) if as_numpy else (y, p)

