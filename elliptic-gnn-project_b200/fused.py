"""The SAGE-ResBN step as ONE explicit kernel sequence (forward and backward), bf16 autocast or fp32.

`/root/reference/src/models/gnn.py:182-194` composes, per hidden layer, PyG SAGEConv -> BatchNorm1d -> ReLU -> dropout
-> `+ res_proj(h_in)`, and autograd replays it backwards op by op.  `ops.py` maps each of those ops to one
`torch.autograd.Function`; this module is the same arithmetic with the op boundaries removed, so that work can move
across them:

  * the layer GEMM  [mean_j h_j | h] . [[W_l, W_r], [0, W_res]]^T  (tcgen05, `egnn_linear_tc`) also produces the
    BatchNorm batch statistics of its own output in its epilogue (no separate pass over z);
  * the gradient of the identity residual joins the root half of the concatenated dgrad GEMM in that GEMM's epilogue
    (no `dx + dres` pass);
  * the layer-0 input `h0 = [x | sin | cos]` (fp32 for the aggregation, bf16 for the GEMM) is a pure function of
    (x, timestep) and is memoised per input version (`StaticInputs`), like the sorted graph views are per
    `edge_index`; `static_inputs(False)` switches the memo off (inputs rewritten in place under a captured graph);
  * gradients are produced in a known order and written straight into caller-provided buffers (`train.TrainStep`
    hands in views of its flat gradient buffer), with no autograd bookkeeping kernels.

`SageResBNFn` wraps the pair as a `torch.autograd.Function` for drop-in use (`loss.backward()`), `train.TrainStep`
calls `forward` / `backward` directly.  Only shapes the tcgen05 kernels take are served (widths a multiple of 8,
hidden <= 128; bf16 operands under bf16 autocast, fp32 operands through the 3xTF32 kernels otherwise -- every
`eval_split` forward is fp32, src/train_gnn.py:248-257); everything else keeps the per-op path of `ops.py`.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional

import torch

from . import _lib, ops
from ._lib import ACT_RELU, check, dt, lib, ptr, stream
from .graph import Graph

_STATIC = True
# weight-gradient GEMMs on a side stream: they depend on dz only, while the chain to the next layer (dgrad GEMM ->
# transposed aggregation -> BatchNorm backward -> cross-rank exchange) is latency-bound on several GPUs -- the side
# stream fills the exchange's waiting time.  "auto": when the step is sharded (a statistics reducer is attached).
# weight-gradient GEMMs (and the weight packing of the forward) on a side stream: they are off the dependency chain of
# the step, so they fill the waiting time of a BatchNorm exchange when sharded and the tails / launch gaps of the
# narrow kernels on one GPU (rec_k8, N=1: 485 -> 465-478 us per graphed step).  False: everything on one stream.
OVERLAP_WGRAD = True
_SIDE = {}


def _side_stream(dev):
    st = _SIDE.get(dev)
    if st is None:
        st = _SIDE[dev] = torch.cuda.Stream(device=dev)
    return st


@contextlib.contextmanager
def static_inputs(enabled: bool):
    """Inside: whether layer-0 input layouts may be memoised per (tensor, version)."""
    global _STATIC
    old, _STATIC = _STATIC, bool(enabled)
    try:
        yield
    finally:
        _STATIC = old


class StaticInputs:
    """Memo of the layer-0 input layout, keyed on storage pointer / shape / in-place version of `x` and `t` and on the
    time table's contents version (a learned table changes every step and is never memoised)."""

    def __init__(self, max_entries: int = 4):
        self._d: Dict[tuple, tuple] = {}
        self.max_entries = max_entries
        self.hits = self.misses = 0

    @staticmethod
    def _key(x, t, table, width):
        k = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x._version, width)
        if t is not None:
            k += (t.data_ptr(), t._version, table.data_ptr(), table._version, tuple(table.shape))
        return k

    def get(self, x, t, table, width, build):
        if not _STATIC or (table is not None and table.requires_grad):
            return build()
        key = self._key(x, t, table, width)
        hit = self._d.get(key)
        if hit is not None:
            self.hits += 1
            return hit[0]
        self.misses += 1
        val = build()
        if len(self._d) >= self.max_entries:
            self._d.pop(next(iter(self._d)))
        self._d[key] = (val, x, t, table)       # keep the key tensors alive: their addresses cannot be recycled
        return val

    def clear(self):
        self._d.clear()


STATIC_INPUTS = StaticInputs()


def supported(net, x: torch.Tensor, bf16: bool) -> bool:
    """Shapes / modes the explicit path serves."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.size(0) < 1024:
        return False
    H = net.convs[0].out_channels
    if H % 16 or H > 128 or net.in_dim > 248:        # 2H and 2K columns per GEMM: N <= 256, K <= 512
        return False
    if any(c.out_channels != H for c in net.convs[:-1]) or net.convs[-1].out_channels not in (1, 2, 4):
        return False
    if net.time_emb is not None and (net.in_dim + 7) // 8 * 8 > 128:
        return False                                  # learned table: the layer-0 dgrad [N, 2K] must fit one GEMM
    return True


class _Layer:
    __slots__ = ("cat", "wcat", "wt", "z", "res_is_input", "has_proj", "mean", "rstd", "kb", "K", "No", "Nr", "p_eff",
                 "y")


class Saved:
    """What the backward needs from the forward."""
    __slots__ = ("layers", "g", "t", "x_cols", "h_last", "wout", "C", "seed", "soff", "n_total", "reducer", "row0",
                 "use_bn", "D", "T")


def _bn_stats_to_mean_rstd(parts, n_parts, F, n_total, bn, reducer, dev):
    L = lib()
    mean = torch.empty(F, dtype=torch.float32, device=dev)
    rstd = torch.empty(F, dtype=torch.float32, device=dev)
    track = bn.track_running_stats and bn.running_mean is not None
    if reducer is None:
        check(L.egnn_bn_finalize_parts(ptr(parts), n_parts, F, float(n_total), float(bn.eps), float(bn.momentum),
                                       ptr(mean), ptr(rstd), ptr(bn.running_mean), ptr(bn.running_var),
                                       ptr(bn.num_batches_tracked) if track else None, stream()))
    elif getattr(reducer, "fused_args", None) is not None and reducer.fused_args(F) is not None:
        # sharded, peer memory: reduce the parts + exchange + finalise in ONE single-block kernel
        check(L.egnn_bn_stats_exchange(ptr(parts), n_parts, F, float(n_total), float(bn.eps), float(bn.momentum),
                                       ptr(mean), ptr(rstd), ptr(bn.running_mean), ptr(bn.running_var),
                                       ptr(bn.num_batches_tracked) if track else None, *reducer.fused_args(F), stream()))
    else:
        if track:
            check(L.egnn_counter_add(ptr(bn.num_batches_tracked), 1, stream()))
        st = torch.empty((2, F), dtype=torch.float64, device=dev)
        check(L.egnn_colstats_reduce(ptr(parts), n_parts, F, ptr(st), stream()))
        reducer.reduce_(st)
        check(L.egnn_bn_finalize(st[0].data_ptr(), st[1].data_ptr(), float(n_total), F, float(bn.eps),
                                 float(bn.momentum), ptr(mean), ptr(rstd), ptr(bn.running_mean), ptr(bn.running_var),
                                 stream()))
    return mean, rstd


def _linear_tc(A, W, out, bias=None, row_div=None, row_div_cols=0, addend=None, add_col0=0, stats=None, stats_cols=0):
    M, K = A.shape
    N = W.size(0)
    L = lib()
    nws = L.egnn_linear_tc_workspace_floats(dt(A), N, K)      # fp32 operands: room for the hi / lo split of W
    ws = torch.empty(nws, dtype=torch.float32, device=A.device) if nws else None
    check(L.egnn_linear_tc(ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(out), dt(out), out.stride(0), M, N, K,
                           ptr(bias), ptr(row_div), int(row_div_cols), 0, ptr(addend),
                           addend.stride(0) if addend is not None else 0, int(add_col0), ptr(stats),
                           int(stats_cols), dt(A), ptr(ws), stream()))
    return out


def forward(net, x: torch.Tensor, g: Graph, t: Optional[torch.Tensor], training: bool, need_grad: bool,
            bf16: bool = True):
    """logits [N, C] fp32 and (when `need_grad`) the Saved state.  Activations in bf16 (`bf16`, autocast) or fp32;
    fp32 statistics / logits either way."""
    L = lib()
    dev = x.device
    N = x.size(0)
    cd = torch.bfloat16 if bf16 else torch.float32
    H = net.convs[0].out_channels
    n_hidden = len(net.convs) - 1
    reducer = net.stats_reducer
    n_total = float(reducer.n_total) if (reducer is not None and reducer.n_total) else float(N)
    drop = None
    if training and net.dropout > 0:
        drop = net.dropout_state(dev)
        drop.advance()
    sv = Saved() if need_grad else None

    # ---- layer-0 input: h0 = [x | time features | 0-pad], fp32 (aggregation input) + bf16 inside [agg | h0]
    K0 = net.in_dim
    K0p = (K0 + 7) // 8 * 8
    use_t = net.time_embed_dim > 0 and t is not None
    table = (net.time_emb.weight if net.time_embed_type == "learned" else net._sin_table) if use_t else None
    if not use_t and x.size(1) != K0:
        raise ValueError(f"expected x of shape [N, {K0}], got {tuple(x.shape)}")
    x = ops._rows(x)

    def build_h0():
        cat0 = torch.empty((N, 2 * K0p), dtype=cd, device=dev)
        tb = table.detach().contiguous().float() if table is not None else None
        if bf16:     # fp32 copy for the aggregation (PyG aggregates the fp32 input in fp32) + bf16 for the GEMM
            h32 = torch.empty((N, K0p), dtype=torch.float32, device=dev)
            o32, ld32, o16 = h32, K0p, cat0[:, K0p:]
        else:        # fp32: h0 lives in the right half of [agg | h0] only
            h32 = cat0[:, K0p:]
            o32, ld32, o16 = h32, 2 * K0p, None
        check(L.egnn_inject_time(ptr(x), ops._ld(x), ptr(t.contiguous()) if use_t else None, ptr(tb),
                                 tb.size(0) if tb is not None else 0, tb.size(1) if tb is not None else 0, ptr(o32),
                                 ptr(o16), ld32, 2 * K0p if o16 is not None else 0, N, x.size(1), stream()))
        return h32, cat0

    h32, cat = STATIC_INPUTS.get(x, t if use_t else None, table, (K0p, bf16), build_h0)

    # Operand packing depends on the parameters only: every layer's [[W_l | W_r], [0 | W_res]] (+ its transpose for the
    # dgrad) and the logits layer's [W_l ; W_r] are written on the side stream while the layer-0 aggregation runs.
    out_conv = net.convs[-1]
    C_out = out_conv.out_channels
    wout = torch.empty((2 * C_out, H), dtype=torch.float32, device=dev)
    no_res = getattr(net, "no_residual", False)      # plain SAGENet (gnn.py:35-53): dropout(relu(conv(h))), no skip
    packed = []
    Kp, Kr = K0p, K0
    for li in range(n_hidden):
        proj = None if no_res else net.res_projs[li]
        has_proj = proj is not None and not isinstance(proj, torch.nn.Identity)
        Nr = H if has_proj else 0
        wcat = torch.empty((H + Nr, 2 * Kp), dtype=cd, device=dev)
        bias = torch.empty(H + Nr, dtype=torch.float32, device=dev)
        want_wt = need_grad and (li > 0 or net.time_emb is not None)
        wt = torch.empty((2 * Kp, H), dtype=cd, device=dev) if want_wt else None
        packed.append((proj, has_proj, Nr, wcat, bias, wt, Kp, Kr))
        Kp, Kr = H, H
    main_st = torch.cuda.current_stream(dev)
    side_st = _side_stream(dev) if (OVERLAP_WGRAD and need_grad) else main_st   # inference: one stream (eager launches)
    if side_st is not main_st:
        ev_fork = torch.cuda.Event()
        ev_fork.record(main_st)
        side_st.wait_event(ev_fork)
    with (torch.cuda.stream(side_st) if side_st is not main_st else contextlib.nullcontext()):   # no stream switch on the eager inference path
        for li, (proj, has_proj, Nr, wcat, bias, wt, Kp_, Kr_) in enumerate(packed):
            conv = net.convs[li]
            check(L.egnn_pack_sage_weights(ptr(conv.lin_l.weight), ptr(conv.lin_r.weight),
                                           ptr(proj.weight) if has_proj else None, ptr(conv.lin_l.bias), H, Nr, Kr_, Kp_,
                                           ptr(wcat), ptr(bias), ptr(wt), dt(wcat), stream()))
        check(L.egnn_concat2_f32(ptr(out_conv.lin_l.weight), C_out * H, ptr(out_conv.lin_r.weight), C_out * H,
                                 ptr(wout), stream()))
    ops.spmm(g, "csr", _lib.SPMM_MEAN, h32, cd, out=cat[:, :K0p])
    if side_st is not main_st:
        main_st.wait_stream(side_st)
    p_out = None

    layers: List[_Layer] = []
    K, Kraw = K0p, K0
    h_in = cat[:, K0p:]
    for li in range(n_hidden):
        conv = net.convs[li]
        # the reference adds `res_projs[li](h_in)` whatever `residual` says (src/models/gnn.py:192; the flag is stored
        # but never read), and so does this path
        proj, has_proj, Nr, wcat, bias, wt, _, _ = packed[li]
        zc = torch.empty((N, H + Nr), dtype=cd, device=dev)
        bn = net.bns[li] if net.use_bn else None
        stats_in_gemm = bn is not None and training and H <= 64
        parts = n_parts = None
        if stats_in_gemm:
            n_parts = int(L.egnn_linear_stats_parts(N))
            parts = torch.empty((n_parts, 2, H), dtype=torch.float32, device=dev)
        _linear_tc(cat, wcat, zc, bias=bias, stats=parts, stats_cols=H if stats_in_gemm else 0)
        z = zc[:, :H]
        res = zc[:, H:] if has_proj else (None if no_res else h_in)
        mean = rstd = None
        if bn is not None:
            if training:
                if stats_in_gemm:
                    mean, rstd = _bn_stats_to_mean_rstd(parts, n_parts, H, n_total, bn, reducer, dev)
                else:
                    if bn.track_running_stats:
                        check(L.egnn_counter_add(ptr(bn.num_batches_tracked), 1, stream()))
                    st = ops.colsum(z, want_sq=True)
                    if reducer is not None:
                        reducer.reduce_(st)
                    mean = torch.empty(H, dtype=torch.float32, device=dev)
                    rstd = torch.empty(H, dtype=torch.float32, device=dev)
                    check(L.egnn_bn_finalize(st[0].data_ptr(), st[1].data_ptr(), n_total, H, float(bn.eps),
                                             float(bn.momentum), ptr(mean), ptr(rstd), ptr(bn.running_mean),
                                             ptr(bn.running_var), stream()))
            else:
                mean, rstd = bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
        p_eff = float(net.dropout) if (training and net.dropout > 0) else 0.0
        last = li == n_hidden - 1
        nxt = torch.empty((N, 2 * H), dtype=cd, device=dev)     # [agg | y]: the next SAGEConv aggregates in place
        y = nxt[:, H:]
        kb = torch.empty((N, H // 4), dtype=torch.uint8, device=dev) if need_grad else None
        # the last hidden layer also emits p = y [W_l ; W_r]^T of the logits layer (project-first), from the same pass
        fold = last and C_out == 2 and H <= 256 and (H // 8) & (H // 8 - 1) == 0
        if fold:
            p_out = torch.empty((N, 2 * C_out), dtype=torch.float32, device=dev)
        check(L.egnn_bn_act_dropout_res_fwd(ptr(z), ptr(res), ptr(y), dt(z), z.stride(0), N, H, ptr(mean), ptr(rstd),
                                            ptr(bn.weight) if bn is not None else None,
                                            ptr(bn.bias) if bn is not None else None, ACT_RELU, p_eff,
                                            drop.seed if drop is not None else 0,
                                            ptr(drop.offset) if drop is not None else None, li, net.row0,
                                            res.stride(0) if res is not None else 0, y.stride(0), ptr(kb),
                                            ptr(wout) if fold else None, ptr(p_out) if fold else None, stream()))
        if need_grad:
            ly = _Layer()
            ly.cat, ly.wcat, ly.wt, ly.z, ly.mean, ly.rstd, ly.kb = cat, wcat, wt, z, mean, rstd, kb
            ly.has_proj, ly.res_is_input = has_proj, not has_proj and not no_res
            ly.K, ly.No, ly.Nr, ly.p_eff, ly.y = K, H, Nr, p_eff, y
            layers.append(ly)
        if not last:
            ops.spmm(g, "csr", _lib.SPMM_MEAN, y, cd, out=nxt[:, :H])
        cat, K, Kraw, h_in = nxt, H, H, y

    # ---- logits layer, project first: p = h [W_l; W_r]^T, out_i = mean_j p_j[:C] + b + p_i[C:]
    C = C_out
    p = p_out
    if p is None:
        p = torch.empty((N, 2 * C), dtype=torch.float32, device=dev)
        check(L.egnn_skinny_project(ptr(h_in), dt(h_in), h_in.stride(0), N, H, ptr(wout), 2 * C, ptr(p), stream()))
    logits = torch.empty((N, C), dtype=torch.float32, device=dev)
    tmp = torch.empty((g.cap, C), dtype=torch.float32, device=dev)
    check(L.egnn_sage_out_fwd(ptr(g.csr_ptr), ptr(g.csr_src), ptr(p), ptr(out_conv.lin_l.bias), C, ptr(logits), N,
                              ptr(tmp), g.cap, stream()))
    if need_grad:
        sv.layers, sv.g, sv.t, sv.x_cols = layers, g, (t if use_t else None), x.size(1)
        sv.h_last, sv.wout, sv.C = h_in, wout, C
        sv.seed, sv.soff = (drop.seed if drop is not None else 0), (drop.offset if drop is not None else None)
        sv.n_total, sv.reducer, sv.row0, sv.use_bn = n_total, reducer, net.row0, net.use_bn
        sv.D = net.time_embed_dim if (use_t and net.time_embed_type == "learned") else 0
        sv.T = net.max_timestep
    return logits, sv


def param_order(net) -> List[torch.nn.Parameter]:
    """The parameters `backward` produces gradients for, in the order it returns them."""
    ps: List[torch.nn.Parameter] = []
    n_hidden = len(net.convs) - 1
    for li in range(n_hidden):
        c = net.convs[li]
        ps += [c.lin_l.weight, c.lin_l.bias, c.lin_r.weight]
        if net.use_bn:
            ps += [net.bns[li].weight, net.bns[li].bias]
        if not getattr(net, "no_residual", False) and not isinstance(net.res_projs[li], torch.nn.Identity):
            ps.append(net.res_projs[li].weight)
    c = net.convs[-1]
    ps += [c.lin_l.weight, c.lin_l.bias, c.lin_r.weight]
    if net.time_emb is not None:
        ps.append(net.time_emb.weight)
    return ps


def backward(net, sv: Saved, dlogits: torch.Tensor, out: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
    """Gradients of every parameter in `param_order(net)`, fp32; the kernels write them straight into `out[i]`
    (contiguous, the parameter's shape -- `train.TrainStep` passes views of its flat gradient buffer) when given.
    `dlogits` fp32 [N, C]."""
    L = lib()
    g = sv.g
    dev = dlogits.device
    ops._ensure_exact()       # fp32 operands: the backward's GEMMs accumulate exactly (ops.set_f32_tc)
    N = dlogits.size(0)
    cd = sv.layers[0].z.dtype
    H = sv.layers[0].No
    C = sv.C
    order = param_order(net)
    if out is None:
        out = [torch.empty(p.shape, dtype=torch.float32, device=dev) for p in order]
    index = {id(p): k for k, p in enumerate(order)}
    dst = lambda param: out[index[id(param)]]
    f32 = dict(dtype=torch.float32, device=dev)

    def copy_into(param, value):
        """small strided fp32 block -> the parameter's gradient buffer"""
        d = dst(param)
        v2 = value if value.dim() == 2 else value.reshape(1, -1)
        d2 = d if d.dim() == 2 else d.reshape(1, -1)
        check(L.egnn_cast(ptr(v2), dt(v2), v2.stride(0) if v2.size(0) > 1 else v2.size(1), ptr(d2), dt(d2),
                          d2.stride(0) if d2.size(0) > 1 else d2.size(1), v2.size(0), v2.size(1), stream()))

    def wgrad(G, X, K, Kraw, d0, d1, G2=None, d2=None):
        """[d0 | d1] = G^T X split at column K, zero-padding columns >= Kraw dropped; d2 = (G2^T X)[:, K:]"""
        No, Kin = G.size(1), X.size(1)
        N2 = G2.size(1) if G2 is not None else 0
        if cd == torch.float32 and not ops.F32_TC_WGRAD:
            # EGNN_F32_TC_WGRAD=0: fp32 weight gradient on the exact FFMA kernel (ops.py)
            def put(src, dstt):
                check(L.egnn_cast(ptr(src), dt(src), src.stride(0), ptr(dstt), dt(dstt), dstt.stride(0), src.size(0),
                                  src.size(1), stream()))
            dW = ops.linear_wgrad(G, X, impl=1)
            put(dW[:, :Kraw], d0)
            if d1 is not None:
                put(dW[:, K:K + Kraw], d1)
            if G2 is not None:
                put(ops.linear_wgrad(G2, X[:, K:], impl=1)[:, :Kraw], d2)
            return
        ws = torch.empty(L.egnn_wgrad_tc_workspace_floats(No + N2, Kin), **f32)
        check(L.egnn_wgrad_tc(ptr(G), G.stride(0), ptr(X), X.stride(0), N, No, Kin, ptr(d0), ptr(d1), K, Kraw,
                              ptr(G2), G2.stride(0) if G2 is not None else 0, N2, ptr(d2), dt(G), ptr(ws), stream()))

    overlap = bool(OVERLAP_WGRAD)
    main_st = torch.cuda.current_stream(dev)
    side_st = _side_stream(dev) if overlap else None
    keep = []          # operands of side-stream kernels stay referenced until the streams join

    # ---- logits layer
    oc = net.convs[-1]
    dlogits = ops._rows(dlogits).contiguous()
    dp = torch.empty((N, 2 * C), **f32)
    tmp = torch.empty((g.cap, C), **f32)
    check(L.egnn_sage_out_bwd(ptr(g.csc_ptr), ptr(g.csc_dst), ptr(g.csr_ptr), ptr(dlogits), dt(dlogits), C, ptr(dp), N,
                              ptr(tmp), g.cap, stream()))
    h = sv.h_last

    def out_wgrad():
        ws = torch.empty(L.egnn_skinny_wgrad_workspace_floats(N, H, 2 * C), **f32)
        check(L.egnn_skinny_wgrad_split(ptr(h), dt(h), h.stride(0), ptr(dp), 2 * C, N, H, ptr(dst(oc.lin_l.weight)),
                                        ptr(dst(oc.lin_r.weight)), ptr(dst(oc.lin_l.bias)), ptr(ws), stream()))

    if overlap:
        ev0 = torch.cuda.Event()
        ev0.record(main_st)
        keep.append((dp, h))
        with torch.cuda.stream(side_st):
            side_st.wait_event(ev0)
            out_wgrad()
    else:
        out_wgrad()
    dy = torch.empty((N, H), dtype=cd, device=dev)
    # dy = dp . [W_l ; W_r] is produced INSIDE the last layer's BatchNorm backward (reduce pass) when that layer has
    # BatchNorm and the shapes fit its 8-column kernel; otherwise by its own pass here
    dy_from_dp = sv.use_bn and C == 2 and (H // 8) & (H // 8 - 1) == 0 and sv.layers[-1].kb is not None
    if not dy_from_dp:
        check(L.egnn_skinny_dgrad(ptr(dp), ptr(sv.wout), 2 * C, ptr(dy), dt(dy), H, N, H, stream()))

    # ---- hidden layers, last to first
    dh0 = None
    for li in range(len(sv.layers) - 1, -1, -1):
        ly = sv.layers[li]
        conv = net.convs[li]
        bn = net.bns[li] if sv.use_bn else None
        z, K = ly.z, ly.K
        dz = torch.empty((N, H), dtype=cd, device=dev)
        dzsum = dst(conv.lin_l.bias)                    # column sums of dz = the conv bias gradient
        ws2 = torch.empty(L.egnn_colreduce_workspace_bytes(H) + 64 * H, dtype=torch.uint8, device=dev)
        if bn is not None:
            fused_sums = ly.kb is not None and (H // 8) & (H // 8 - 1) == 0 and H <= 256   # the 8-column reduce kernel
            sg = torch.empty((2, H), dtype=torch.float64, device=dev)
            wsr = torch.empty(L.egnn_colreduce_workspace_bytes(H), dtype=torch.uint8, device=dev)
            xargs = None
            if fused_sums and getattr(sv.reducer, "fused_args", None) is not None:
                xargs = sv.reducer.fused_args(H)
            check(L.egnn_bn_act_dropout_bwd_reduce(ptr(dy), ptr(z), dt(z), H, N, H, ptr(ly.mean), ptr(ly.rstd),
                                                   ptr(bn.weight), ptr(bn.bias), ACT_RELU, ly.p_eff, sv.seed,
                                                   ptr(sv.soff), li, sv.row0,
                                                   sg[0].data_ptr() if xargs is None else None,
                                                   sg[1].data_ptr() if xargs is None else None,
                                                   ptr(wsr), z.stride(0), ptr(ly.kb),
                                                   ptr(dp) if (dy_from_dp and li == len(sv.layers) - 1) else None,
                                                   ptr(sv.wout) if (dy_from_dp and li == len(sv.layers) - 1) else None,
                                                   ptr(dst(bn.bias)) if (fused_sums and sv.reducer is None) else None,
                                                   ptr(dst(bn.weight)) if (fused_sums and sv.reducer is None) else None,
                                                   stream()))
            if xargs is not None:      # partial rows -> all ranks' [sum g | sum g*xhat] (+ d beta / d gamma) in one kernel
                check(L.egnn_bn_bwd_sums_exchange(ptr(wsr), int(L.egnn_bn_bwd_reduce_parts(N, H)), H, ptr(sg),
                                                  ptr(dst(bn.bias)), ptr(dst(bn.weight)), *xargs, stream()))
            elif sv.reducer is not None:
                # d beta / d gamma = THIS rank's share (the gradient all-reduce adds the shares), then the exchange
                check(L.egnn_f64_to_f32(sg[0].data_ptr(), ptr(dst(bn.bias)), H, stream()))
                check(L.egnn_f64_to_f32(sg[1].data_ptr(), ptr(dst(bn.weight)), H, stream()))
                sv.reducer.reduce_(sg)
            check(L.egnn_bn_act_dropout_bwd_apply(ptr(dy), ptr(z), ptr(dz), dt(z), H, N, H, ptr(ly.mean), ptr(ly.rstd),
                                                  ptr(bn.weight), ptr(bn.bias), ACT_RELU, ly.p_eff, sv.seed,
                                                  ptr(sv.soff), li, sv.row0, sg[0].data_ptr(), sg[1].data_ptr(),
                                                  sv.n_total, ptr(dzsum), ptr(ws2), z.stride(0), ptr(ly.kb), stream()))
            if not fused_sums and sv.reducer is None:
                check(L.egnn_f64_to_f32(sg[0].data_ptr(), ptr(dst(bn.bias)), H, stream()))      # d beta  = sum g
                check(L.egnn_f64_to_f32(sg[1].data_ptr(), ptr(dst(bn.weight)), H, stream()))    # d gamma = sum g * xhat
        else:
            check(L.egnn_bn_act_dropout_bwd_apply(ptr(dy), ptr(z), ptr(dz), dt(z), H, N, H, None, None, None, None,
                                                  ACT_RELU, ly.p_eff, sv.seed, ptr(sv.soff), li, sv.row0, None, None,
                                                  1.0, ptr(dzsum), ptr(ws2), z.stride(0), ptr(ly.kb), stream()))
        # weight gradients: dz^T [m | h]  (and dres^T h for a projected residual; dres = dy)
        cat = ly.cat
        Kraw = conv.in_channels

        def layer_wgrads():
            if 2 * K <= 384 and ly.has_proj and H % 64 == 0 and 2 * H <= 128 and cd == torch.bfloat16:
                # one pass over [m | h]: [dz | dy]^T [m | h] -> d lin_l, d lin_r, and (dy x root half) d res_proj
                wgrad(dz, cat, K, Kraw, dst(conv.lin_l.weight), dst(conv.lin_r.weight), G2=dy,
                      d2=dst(net.res_projs[li].weight))
            else:
                if 2 * K <= 384:
                    wgrad(dz, cat, K, Kraw, dst(conv.lin_l.weight), dst(conv.lin_r.weight))
                else:
                    wgrad(dz, cat[:, :K], K, Kraw, dst(conv.lin_l.weight), None)
                    wgrad(dz, cat[:, K:], K, Kraw, dst(conv.lin_r.weight), None)
                if ly.has_proj:
                    wgrad(dy, cat[:, K:], K, Kraw, dst(net.res_projs[li].weight), None)

        if overlap and li > 0:
            ev = torch.cuda.Event()
            ev.record(main_st)
            keep.append((dz, dy, cat))
            with torch.cuda.stream(side_st):
                side_st.wait_event(ev)
                layer_wgrads()
        else:
            layer_wgrads()
        need_dh = li > 0 or sv.D > 0
        if not need_dh:
            break
        # [dm / in-degree | dx_root (+ dy of an identity residual)] = dz . [W_l | W_r]; dh = dx_root + A^T (dm / deg)
        o2 = torch.empty((N, 2 * K), dtype=cd, device=dev)      # ly.wt = [W_l | W_r]^T, packed by the forward
        _linear_tc(dz, ly.wt, o2, row_div=g.csr_ptr, row_div_cols=K, addend=dy if ly.res_is_input else None,
                   add_col0=K)
        if ly.has_proj:
            ops.linear_dgrad(dy, ly.wcat[H:, K:], out=o2[:, K:], accumulate=True)
        dh = ops.spmm(g, "csc", _lib.SPMM_SUM, o2[:, :K], cd, addend=o2[:, K:])
        if li == 0:
            dh0 = dh
        dy = dh
    if overlap:
        main_st.wait_stream(side_st)
        keep.clear()
    if sv.D > 0:
        wse = torch.empty(L.egnn_embed_grad_workspace_bytes(N, sv.T, sv.D), dtype=torch.uint8, device=dev)
        check(L.egnn_embed_grad(ptr(dh0), dt(dh0), dh0.stride(0), sv.x_cols, sv.D, ptr(sv.t.contiguous()), sv.T, N,
                                ptr(dst(net.time_emb.weight)), ptr(wse), stream()))
    return out


class SageResBNFn(torch.autograd.Function):
    """Drop-in autograd wrapper: logits = SageResBNFn.apply(net, x, g, t, bf16, *param_order(net))."""

    @staticmethod
    def forward(ctx, net, x, g, t, bf16, *params):
        need = any(ctx.needs_input_grad[5:])
        logits, sv = forward(net, x, g, t, net.training, need, bf16)
        ctx.net, ctx.sv = net, sv
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        grads = backward(ctx.net, ctx.sv, dlogits.float())
        ctx.sv = None
        return (None, None, None, None, None) + tuple(grads)
