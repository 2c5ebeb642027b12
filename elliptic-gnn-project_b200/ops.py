"""torch.autograd.Function wrappers over the C-ABI kernels.

Each Function is the B200 replacement of one PyG / torch op sequence on the reference's
hot path (`/root/reference/src/models/gnn.py`); forward and backward both run hand-written
sm_100a kernels from libegnn_b200.so.  Nothing here falls back to CPU or to PyTorch ops
for the heavy work; torch is used for allocation (stream-ordered caching allocator),
autograd bookkeeping and a few O(F)-sized conversions.
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_ELU, ACT_NONE, ACT_RELU, BF16, F32, check, dt, lib, ptr, stream
from .graph import Graph

_TD = {F32: torch.float32, BF16: torch.bfloat16}


_FP16_WARNED = False


def _fp16_policy() -> str:
    """`EGNN_FP16_AUTOCAST`: "fp32" (default) computes an fp16-autocast region in fp32; "raise" rejects it."""
    p = os.environ.get("EGNN_FP16_AUTOCAST", "fp32").strip().lower()
    if p not in ("fp32", "raise"):
        raise ValueError(f"EGNN_FP16_AUTOCAST must be 'fp32' or 'raise', got {p!r}")
    return p


def amp_bf16() -> bool:
    """True when the caller runs us under torch.autocast(device_type='cuda', dtype=bfloat16).

    The reference's `amp: true` is fp16 autocast + GradScaler (`src/train_gnn.py:36-47,202-207`).  There are no fp16
    kernels here: under fp16 autocast the convs and nets compute in fp32 (tensor cores, 3xTF32 with exact
    accumulation) -- every value the reference rounds to fp16 is the fp32 value computed here, so the results lie
    inside the reference's own rounding error -- and emit fp32, which `GradScaler.scale / unscale_ / step` handle
    like any fp32 gradient.  It costs the bf16 path's speed (use autocast(dtype=torch.bfloat16) for that), hence the
    one-time warning; `EGNN_FP16_AUTOCAST=raise` turns the widening into an error."""
    global _FP16_WARNED
    if not torch.is_autocast_enabled("cuda"):
        return False
    d = torch.get_autocast_dtype("cuda")
    if d == torch.bfloat16:
        return True
    if d == torch.float16 and _fp16_policy() == "fp32":
        if not _FP16_WARNED:
            _FP16_WARNED = True
            warnings.warn("egnn_b200: fp16 autocast region computed in fp32 (no fp16 kernels; results are at least as "
                          "accurate as fp16).  Use torch.autocast('cuda', dtype=torch.bfloat16) for the fast path.",
                          RuntimeWarning, stacklevel=3)
        return False
    raise RuntimeError(
        "egnn_b200 implements bf16 autocast only; the reference's `amp: true` defaults to fp16 + "
        "GradScaler (src/train_gnn.py:36-47).  Use torch.autocast('cuda', dtype=torch.bfloat16).")


def widen_fp16(x: torch.Tensor) -> torch.Tensor:
    """fp16 activations (what a caller's own `nn.Linear` emits under fp16 autocast) enter the kernels as fp32: the
    conversion is exact and differentiable (the upstream module receives an fp16 gradient, as autocast would give it)."""
    if x.dtype != torch.float16:
        return x
    if _fp16_policy() == "raise":
        raise TypeError("egnn_b200 kernels take float32 or bfloat16 tensors, got torch.float16 "
                        "(EGNN_FP16_AUTOCAST=raise)")
    return x.float()


def _rows(t: torch.Tensor) -> torch.Tensor:
    """2-D, unit stride along features (row stride may exceed the width)."""
    if t.dim() != 2:
        raise ValueError("expected a 2-D [rows, features] tensor")
    if t.size(1) > 1 and t.stride(1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), t.stride(0))


# ------------------------------------------------------------------------------ raw ops --
def to_compute(x: torch.Tensor, cd: torch.dtype) -> torch.Tensor:
    """x in the compute dtype: x itself, its pre-computed twin (an op that produced x may have
    written a bf16 copy in the same pass, see InjectTimeFn), or a cast."""
    if x.dtype == cd:
        return x
    tw = getattr(x, "_egnn_twin", None)
    if tw is not None and tw.dtype == cd and tw.shape == x.shape:
        return tw
    return cast(x, cd)


def alloc_cat(n_rows: int, width: int, dtype: torch.dtype, device):
    """An activation buffer with room for its own aggregation: buf = [agg | h], [n_rows, 2*width].
    Returns (buf, h) where h = buf[:, width:] carries `_egnn_cat = buf`; the next SAGEConv writes
    mean_j h_j into buf[:, :width] and runs ONE GEMM over the concatenated operand."""
    buf = torch.empty((n_rows, 2 * width), dtype=dtype, device=device)
    h = buf[:, width:]
    h._egnn_cat = buf
    return buf, h


def _lean4(F: int, *tensors) -> bool:
    """Shapes served by the 4-column streaming kernels (which also take per-tensor row strides):
    F/4 a power of two <= 256, 4-element aligned pointers and row strides."""
    g = F // 4
    if F % 4 or g > 256 or g & (g - 1):
        return False
    for t in tensors:
        if t is None:
            continue
        if t.stride(-1) != 1 or t.data_ptr() % (4 * t.element_size()) or (t.size(0) > 1 and t.stride(0) % 4):
            return False
    return True


def pad_features(x: torch.Tensor, width: int, want_twin: bool, twin_read_only: bool = False):
    """Raw fp32 node features zero-padded to `width` columns (a multiple of 8: 16-byte rows for the
    vectorised gathers and the TMA descriptors; e.g. the reference's 167 = 166 + scalar time,
    `src/train_gnn.py:314-317`), plus -- under bf16 autocast -- the bf16 copy inside an [agg | h]
    buffer, both written by one pass of the time-injection kernel with an empty time table.

    The result is memoised on the input tensor object per version of its data (static inputs: the reference moves the
    graph to the device once) unless the caller will WRITE into the twin's buffer: a SAGEConv aggregates into the
    [agg | h] slot and saves it for its backward, so it always gets a fresh one; GCN / GAT only read the twin
    (`twin_read_only`)."""
    from . import fused                          # fused.static_inputs(False): inputs rewritten in place under a graph
    share = fused._STATIC and (not want_twin or twin_read_only)
    key = (x._version, x.data_ptr(), width, bool(want_twin))
    memo = getattr(x, "_egnn_padded", None) if share else None
    if memo is not None and memo[0] == key:
        return memo[1]
    src = x
    x = _rows(x)
    N, F = x.shape
    out = torch.empty((N, width), dtype=torch.float32, device=x.device)
    twin = alloc_cat(N, width, torch.bfloat16, x.device)[1] if want_twin else None
    check(lib().egnn_inject_time(ptr(x), _ld(x), None, None, 0, 0, ptr(out), ptr(twin), width,
                                 _ld(twin) if twin is not None else 0, N, F, stream()))
    if twin is not None:
        out._egnn_twin = twin
    if share and not src.requires_grad and not torch.cuda.is_current_stream_capturing():
        src._egnn_padded = (key, out)
    return out


def _pad_cols(w: torch.Tensor, pad: int) -> torch.Tensor:
    return torch.cat([w, w.new_zeros((w.size(0), pad))], dim=1) if pad else w


def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    x = _rows(x)
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(lib().egnn_cast(ptr(x), dt(x), _ld(x), ptr(out), dt(out), out.size(1), x.size(0), x.size(1),
                          stream()))
    return out


def spmm(g: Graph, view: str, mode: int, x: torch.Tensor, out_dtype: torch.dtype, *, bias=None,
         act: int = ACT_NONE, out: Optional[torch.Tensor] = None, accumulate: bool = False,
         addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    x = _rows(x)
    n = g.n_nodes
    if x.size(0) != n:
        raise ValueError(f"feature matrix has {x.size(0)} rows, graph has {n} nodes")
    if out is None:
        out = torch.empty((n, x.size(1)), dtype=out_dtype, device=x.device)
    if view == "csr":
        p, c, w, long_rows, vi, nbr, order = g.csr_ptr, g.csr_src, g.w_csr, g.csr_long, 0, g.csc_ptr, g.csr_order
        part = g.csr_part
    else:
        p, c, w, long_rows, vi, nbr, order = g.csc_ptr, g.csc_dst, g.w_csc, g.csc_long, 1, g.csr_ptr, g.csc_order
        part = g.csc_part
    n_long = g.info.data_ptr() + 4 * (2 + vi)
    check(lib().egnn_spmm(mode, ptr(p), ptr(c), ptr(w) if mode == _lib.SPMM_WEIGHTED else None,
                          ptr(nbr) if mode == _lib.SPMM_DIV_NBR else None, ptr(long_rows), n_long, ptr(order),
                          ptr(part), g.n_tasks if part is not None else 0, ptr(x),
                          dt(x), _ld(x), ptr(out), dt(out), _ld(out), n, x.size(1), ptr(bias), act,
                          int(accumulate), ptr(addend), _ld(addend) if addend is not None else 0, stream()))
    return out


GEMM_IMPL = 0  # 0 auto (tcgen05 for bf16 where supported), 1 force SIMT, 2 force tcgen05

# fp32 operands on the tensor cores (3xTF32 with EXACT accumulation, csrc/gemm_tcgen05.cu).  The tensor core adds into
# its fp32 accumulator with truncation, a bias that does not average out (round 2, first version: parameter gradients at
# ~1e-4 of the oracle's through three layers of forward + backward).  The kernels therefore take every 8-wide k-step
# product out of the tensor core in a fresh TMEM buffer and add it with IEEE round-to-nearest on the CUDA cores (forward /
# dgrad, `kExact`), and split the weight gradient's node axis into phases of <= 132 tensor-core additions combined with
# IEEE adds.  Measured at full size against the CPU oracle (profiles/r02/f32_tc_probe.txt): gcn.yaml / sage.yaml
# gradients within 2e-6 (FFMA: 4e-7), sage.yaml fp32 1.60 -> 0.66 ms, gcn.yaml 1.75 -> 1.07 ms, rec_k8 fp32 2.72 -> 1.29
# ms per step.  EGNN_F32_TC_TRAIN=0 / EGNN_F32_TC_WGRAD=0 (or the module attributes) put fp32 TRAINING back on the exact
# FFMA kernels (gemm_simt.cu); no-grad forwards always use the tensor cores.
import os as _os

F32_TC_TRAIN = _os.environ.get("EGNN_F32_TC_TRAIN", "1") == "1"
F32_TC_WGRAD = _os.environ.get("EGNN_F32_TC_WGRAD", "1") == "1"
_F32_TC = False
_F32_TC_EXACT = True      # the library's default


def set_f32_tc(grad: Optional[bool] = None):
    """Called by the conv / net modules on entry (where autograd's mode is still visible) and by the explicit train
    step: fp32 GEMMs of this forward (and of its backward) use the tensor cores unless a gradient will be taken AND the
    caller opted out; they accumulate exactly iff a gradient will be taken."""
    global _F32_TC, _F32_TC_EXACT
    if grad is None:
        grad = torch.is_grad_enabled()
    _F32_TC = F32_TC_TRAIN or not grad
    if grad != _F32_TC_EXACT:        # exact accumulation for anything a gradient is taken of (include/egnn_b200.h)
        lib().egnn_set_f32_tc_exact(int(grad))
        _F32_TC_EXACT = grad
    return _F32_TC


def _ensure_exact():
    """Input / weight gradients are only ever computed for a gradient: exact accumulation, whatever a no-grad forward
    that ran between this graph's forward and its backward left the library in."""
    global _F32_TC_EXACT
    if not _F32_TC_EXACT:
        lib().egnn_set_f32_tc_exact(1)
        _F32_TC_EXACT = True


def _gemm(A, a_sm, a_sk, B, b_sk, b_sn, C, M, N, K, bias, accumulate, split_k=1, impl=None, row_div=None,
          need_ws=False, row_div_cols=0):
    L = lib()
    both_f32 = A.dtype == torch.float32 and B.dtype == torch.float32
    f32_tc = both_f32 and _F32_TC and a_sk == 1 and b_sk == 1 and M >= 1024
    if both_f32 and not _F32_TC and impl is None and GEMM_IMPL == 0:
        impl = 1                     # exact fp32: FFMA kernels
    nws = L.egnn_gemm_workspace_floats(M, N, K, split_k) if (need_ws or split_k > 1 or M <= 8 or f32_tc) else 0
    ws = torch.empty(nws, dtype=torch.float32, device=C.device) if nws else None
    check(L.egnn_gemm(ptr(A), dt(A), a_sm, a_sk, ptr(B), dt(B), b_sk, b_sn, ptr(C), dt(C), _ld(C), M, N, K,
                      ptr(bias), ptr(row_div), int(row_div_cols), int(accumulate), split_k, ptr(ws),
                      GEMM_IMPL if impl is None else impl, stream()))


def linear_fwd(x, W, bias=None, out=None, accumulate=False, out_dtype=None, impl=None):
    """out[M,N] (+)= x[M,K] @ W[N,K]^T (+ bias)."""
    x, W = _rows(x), _rows(W)
    M, K = x.shape
    N = W.size(0)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or x.dtype, device=x.device)
    _gemm(x, _ld(x), 1, W, 1, _ld(W), out, M, N, K, bias, accumulate, impl=impl)
    return out


def linear_dgrad(g, W, out=None, accumulate=False, out_dtype=None, row_div=None, impl=None, row_div_cols=0):
    """out[M,K] (+)= g[M,N] @ W[N,K]; `row_div` = CSR row pointer whose row counts divide the rows
    of the result (SAGE mean backward: dsum = dm / cnt, fused into the GEMM epilogue)."""
    g, W = _rows(g), _rows(W)
    M, N = g.shape
    K = W.size(1)
    if g.dtype == torch.float32:
        _ensure_exact()
    if out is None:
        out = torch.empty((M, K), dtype=out_dtype or g.dtype, device=g.device)
    if g.dtype == W.dtype and g.dtype in (torch.bfloat16, torch.float32):
        Wt = W.t().contiguous()  # [K, N]: contraction-contiguous B operand (tcgen05 kernel / fp32 128-bit-load kernel)
        _gemm(g, _ld(g), 1, Wt, 1, _ld(Wt), out, M, K, N, None, accumulate, row_div=row_div, impl=impl,
              row_div_cols=row_div_cols)
    else:
        _gemm(g, _ld(g), 1, W, _ld(W), 1, out, M, K, N, None, accumulate, row_div=row_div, impl=impl,
              row_div_cols=row_div_cols)
    return out


def linear_wgrad(g, x, impl=None):
    """dW[N,K] = g[M,N]^T @ x[M,K], fp32, deterministic split over the node axis."""
    g, x = _rows(g), _rows(x)
    M, N = g.shape
    K = x.size(1)
    if g.dtype == torch.float32:
        _ensure_exact()
    out = torch.empty((N, K), dtype=torch.float32, device=g.device)
    tiles = -(-N // 128) * -(-K // 64)
    split = max(1, min(-(-M // 512), -(-2 * 148 // tiles)))
    if impl is None and g.dtype == torch.float32 and x.dtype == torch.float32 and not F32_TC_WGRAD:
        impl = 1
    _gemm(g, 1, _ld(g), x, _ld(x), 1, out, N, K, M, None, False, split_k=split, impl=impl, need_ws=True)
    return out


def colsum(a: torch.Tensor, want_sq: bool = False) -> torch.Tensor:
    """double [F] (or [2,F] with sum of squares): deterministic column sums over rows."""
    a = _rows(a)
    F = a.size(1)
    L = lib()
    out = torch.empty((2 if want_sq else 1, F), dtype=torch.float64, device=a.device)
    ws = torch.empty(L.egnn_colreduce_workspace_bytes(F), dtype=torch.uint8, device=a.device)
    check(L.egnn_colreduce(ptr(a), dt(a), _ld(a), a.size(0), F, out[0].data_ptr(),
                           out[1].data_ptr() if want_sq else None, ptr(ws), stream()))
    return out if want_sq else out[0]


def dropout_mask(n_rows: int, n_feat: int, p: float, seed: int, layer: int, row0: int = 0,
                 device="cuda", seed_off: Optional[torch.Tensor] = None) -> torch.Tensor:
    m = torch.empty((n_rows, n_feat), dtype=torch.uint8, device=device)
    check(lib().egnn_dropout_mask(ptr(m), n_rows, n_feat, float(p), int(seed), ptr(seed_off), int(layer),
                                  int(row0), stream()))
    return m


# Column sums of a gradient matrix produced as a by-product of the kernel that wrote it (the
# BatchNorm backward), handed to the consumer (the conv's bias gradient) without another pass
# over the matrix.  The sums ride on the gradient tensor OBJECT itself (autograd hands the very
# tensor a backward returned to the next node when it is that node's only gradient); a tensor
# that autograd re-created (accumulated gradients, hooks) simply lacks the attribute and the
# consumer recomputes the sums -- no global state, nothing keyed on recyclable addresses.
def _publish_colsum(t: torch.Tensor, s: torch.Tensor):
    t._egnn_colsum = (t._version, s)


def _take_colsum(t: torch.Tensor) -> Optional[torch.Tensor]:
    side = getattr(t, "_egnn_colsum", None)
    if side is None or side[0] != t._version or side[1].numel() != t.size(-1):
        return None
    return side[1]


# --------------------------------------------------------------------------- SAGEConv ----
class SageConvFn(torch.autograd.Function):
    """PyG SAGEConv(mean, root_weight, bias): lin_l(mean_j x_j) + lin_r(x)  (SURVEY.md A.2), optionally
    with the residual projection res = x W_res^T of SAGEResBNNet (gnn.py:141-144,192) folded into
    the same GEMM.

    bf16 path with a concatenated activation buffer (`alloc_cat`): the mean aggregation is written
    next to x, so the layer is ONE tensor-core GEMM  [m | x] . [[W_l, W_r], [0, W_res]]^T (+ bias),
    the weight gradient one GEMM dz^T [m | x], and the input gradient one GEMM
    dz . [W_l | W_r] -> [dm/deg | dx_root] followed by the transposed aggregation accumulating
    into its right half."""

    @staticmethod
    def forward(ctx, x, w_l, b_l, w_r, w_res, g: Graph, bf16: bool):
        cd = torch.bfloat16 if bf16 else torch.float32
        K0 = x.size(1)
        pad = (-K0) % 8 if x.dtype == torch.float32 else 0
        if pad:   # e.g. 167 raw columns: pad once so that gathers and GEMM operands have 16-byte rows
            x = pad_features(x, K0 + pad, want_twin=bf16)
        ctx.pad = pad
        xg = to_compute(x, cd)
        x = _rows(x)
        K, No = x.size(1), w_l.size(0)
        cat = getattr(xg, "_egnn_cat", None) if bf16 else None
        if cat is not None and getattr(xg, "_egnn_cat_used", False):
            cat = None          # the aggregation slot already serves another consumer of this tensor
        ctx.g, ctx.x_dtype, ctx.dims = g, x.dtype, (K, No, 0 if w_res is None else w_res.size(0))
        ctx.has_res = w_res is not None
        if cat is not None and cat.dtype == cd and cat.size(1) == 2 * K:
            xg._egnn_cat_used = True
            spmm(g, "csr", _lib.SPMM_MEAN, x, cd, out=cat[:, :K])
            Nr = 0 if w_res is None else w_res.size(0)
            wcat = torch.empty((No + Nr, 2 * K), dtype=cd, device=x.device)     # [[W_l | W_r], [0 | W_res]]
            bias = torch.empty(No + Nr, dtype=torch.float32, device=x.device)
            check(lib().egnn_pack_sage_weights(ptr(w_l.contiguous()), ptr(w_r.contiguous()),
                                               ptr(w_res.contiguous()) if w_res is not None else None,
                                               ptr(b_l.contiguous()) if b_l is not None else None, No, Nr, K0, K,
                                               ptr(wcat), ptr(bias), None, BF16, stream()))
            zc = linear_fwd(cat, wcat, bias=bias, out_dtype=cd)
            ctx.cat_path = True
            ctx.save_for_backward(cat, wcat)
            if w_res is None:
                return zc
            return zc[:, :No], zc[:, No:]
        if pad:
            w_l, w_r = _pad_cols(w_l, pad), _pad_cols(w_r, pad)
            w_res = _pad_cols(w_res, pad) if w_res is not None else None
        m = spmm(g, "csr", _lib.SPMM_MEAN, x, cd)
        wl = w_l if w_l.dtype == cd else cast(w_l, cd)
        wr = w_r if w_r.dtype == cd else cast(w_r, cd)
        z = linear_fwd(m, wl, bias=b_l, out_dtype=cd)
        linear_fwd(xg, wr, out=z, accumulate=True)
        ctx.cat_path = False
        if w_res is None:
            ctx.save_for_backward(m, xg, wl, wr)
            return z
        wres = w_res if w_res.dtype == cd else cast(w_res, cd)
        ctx.save_for_backward(m, xg, wl, wr, wres)
        return z, linear_fwd(xg, wres, out_dtype=cd)

    @staticmethod
    def backward(ctx, dz, dres=None):
        g = ctx.g
        K, No, Nr = ctx.dims
        dz = _rows(dz)
        db = _take_colsum(dz)
        dwres = None
        if ctx.cat_path:
            cat, wcat = ctx.saved_tensors
            cd = cat.dtype
            if dz.dtype != cd:
                dz = cast(dz, cd)
            if 2 * K <= 384:
                dwc = linear_wgrad(dz, cat)
                dwl, dwr = dwc[:, :K], dwc[:, K:]
            else:
                dwl, dwr = linear_wgrad(dz, cat[:, :K]), linear_wgrad(dz, cat[:, K:])
            if ctx.has_res:
                dres = _rows(dres)
                if dres.dtype != cd:
                    dres = cast(dres, cd)
                dwres = linear_wgrad(dres, cat[:, K:])
            if db is None:
                db = colsum(dz).float()
            dx = None
            if ctx.needs_input_grad[0]:
                # [dm / in-degree | dx_root] = dz . [W_l | W_r]; then dx = dx_root + A^T (dm / deg)
                out = linear_dgrad(dz, wcat[:No], row_div=g.csr_ptr, row_div_cols=K)
                if ctx.has_res:
                    linear_dgrad(dres, wcat[No:, K:], out=out[:, K:], accumulate=True)
                # dx (contiguous) = dx_root + A^T (dm / deg)
                dx = spmm(g, "csc", _lib.SPMM_SUM, out[:, :K], cd, addend=out[:, K:])
                if dx.dtype != ctx.x_dtype:
                    dx = cast(dx, ctx.x_dtype)
            return SageConvFn._strip(ctx.pad, dx, dwl, db, dwr, dwres)
        if ctx.has_res:
            m, xg, wl, wr, wres = ctx.saved_tensors
        else:
            m, xg, wl, wr = ctx.saved_tensors
        if dz.dtype != m.dtype:
            dz = cast(dz, m.dtype)
        dwl = linear_wgrad(dz, m)
        dwr = linear_wgrad(dz, xg)
        if db is None:
            db = colsum(dz).float()
        if ctx.has_res:
            dres = _rows(dres)
            if dres.dtype != m.dtype:
                dres = cast(dres, m.dtype)
            dwres = linear_wgrad(dres, xg)
        dx = None
        if ctx.needs_input_grad[0]:
            dm = linear_dgrad(dz, wl, row_div=g.csr_ptr)   # dm / in-degree, fused in the GEMM epilogue
            dx = linear_dgrad(dz, wr)
            if ctx.has_res:
                linear_dgrad(dres, wres, out=dx, accumulate=True)
            spmm(g, "csc", _lib.SPMM_SUM, dm, dx.dtype, out=dx, accumulate=True)
            if dx.dtype != ctx.x_dtype:
                dx = cast(dx, ctx.x_dtype)
        return SageConvFn._strip(ctx.pad, dx, dwl, db, dwr, dwres)

    @staticmethod
    def _strip(pad, dx, dwl, db, dwr, dwres):
        if pad:
            cut = lambda t: None if t is None else t[:, :t.size(1) - pad]
            dx, dwl, dwr, dwres = cut(dx), cut(dwl), cut(dwr), cut(dwres)
        return dx, dwl, db, dwr, dwres, None, None


class SageOutFn(torch.autograd.Function):
    """SAGEConv with <= 4 output channels (the `hidden -> 2` logits layer, gnn.py:44,128),
    evaluated project-first: p = h [W_l; W_r]^T, out = (mean_j p_j[:C] + b) + p_i[C:]  -- three
    passes over h for forward + backward instead of ten (csrc/sage_out.cu).  fp32 weights, fp32
    accumulation and fp32 logits whatever the activation dtype."""

    @staticmethod
    def forward(ctx, x, w_l, b_l, w_r, g: Graph):
        x = _rows(x)
        N, K = x.shape
        C = w_l.size(0)
        L = lib()
        wcat = torch.cat([w_l.detach().float(), w_r.detach().float()], dim=0).contiguous()  # [2C, K]
        p = torch.empty((N, 2 * C), dtype=torch.float32, device=x.device)
        check(L.egnn_skinny_project(ptr(x), dt(x), _ld(x), N, K, ptr(wcat), 2 * C, ptr(p), stream()))
        out = torch.empty((N, C), dtype=torch.float32, device=x.device)
        bias = b_l.detach().float().contiguous() if b_l is not None else None
        tmp = torch.empty((g.cap, C), dtype=torch.float32, device=x.device)
        check(L.egnn_sage_out_fwd(ptr(g.csr_ptr), ptr(g.csr_src), ptr(p), ptr(bias), C, ptr(out), N, ptr(tmp),
                                  g.cap, stream()))
        ctx.g, ctx.C = g, C
        ctx.save_for_backward(x, wcat)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, wcat = ctx.saved_tensors
        g, C = ctx.g, ctx.C
        N, K = x.shape
        L = lib()
        dout = _rows(dout).contiguous()
        dp = torch.empty((N, 2 * C), dtype=torch.float32, device=x.device)
        tmp = torch.empty((g.cap, C), dtype=torch.float32, device=x.device)
        check(L.egnn_sage_out_bwd(ptr(g.csc_ptr), ptr(g.csc_dst), ptr(g.csr_ptr), ptr(dout), dt(dout), C, ptr(dp),
                                  N, ptr(tmp), g.cap, stream()))
        dw = torch.empty((2 * C, K), dtype=torch.float32, device=x.device)
        dsum = torch.empty(2 * C, dtype=torch.float32, device=x.device)
        ws = torch.empty(L.egnn_skinny_wgrad_workspace_floats(N, K, 2 * C), dtype=torch.float32, device=x.device)
        check(L.egnn_skinny_wgrad(ptr(x), dt(x), _ld(x), ptr(dp), 2 * C, N, K, ptr(dw), ptr(dsum), ptr(ws),
                                  stream()))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((N, K), dtype=x.dtype, device=x.device)
            check(L.egnn_skinny_dgrad(ptr(dp), ptr(wcat), 2 * C, ptr(dx), dt(dx), K, N, K, stream()))
        return dx, dw[:C], dsum[C:], dw[C:], None


def sage_out_supported(x: torch.Tensor, out_channels: int) -> bool:
    K = x.size(1)
    return (out_channels in (1, 2, 4) and K % 8 == 0 and K <= 1024 and x.dtype in (torch.float32, torch.bfloat16)
            and x.stride(-1) == 1 and x.data_ptr() % 32 == 0 and (x.size(0) <= 1 or x.stride(0) % 8 == 0))


# ---------------------------------------------------------------------------- GCNConv ----
class GcnConvFn(torch.autograd.Function):
    """PyG GCNConv: D^-1/2 (A+I) D^-1/2 (x W^T) + b on the self-loop graph."""

    @staticmethod
    def forward(ctx, x, w, b, g: Graph, bf16: bool, out_bf16: bool = False):
        cd = torch.bfloat16 if bf16 else torch.float32
        x = _rows(x)
        pad = (-x.size(1)) % 8 if x.dtype == torch.float32 else 0   # 16-byte rows: vectorised / TMA operand loads
        if pad:
            x, w = pad_features(x, x.size(1) + pad, want_twin=bf16, twin_read_only=True), _pad_cols(w, pad)
        ctx.pad = pad
        xg = to_compute(x, cd)
        wc = w if w.dtype == cd else cast(w, cd)
        h = linear_fwd(xg, wc, out_dtype=cd)
        # PyG returns fp32 here even under autocast (fp32 weights promote the aggregation); a hidden layer of OUR
        # nets may ask for the bf16 rounding the next Linear would apply anyway (out_bf16), which halves the
        # activation traffic and removes two cast passes per layer
        out = spmm(g, "csr", _lib.SPMM_WEIGHTED, h, torch.bfloat16 if (out_bf16 and bf16) else torch.float32, bias=b)
        ctx.g, ctx.x_dtype = g, x.dtype
        ctx.save_for_backward(xg, wc)
        return out

    @staticmethod
    def backward(ctx, dout):
        xg, wc = ctx.saved_tensors
        g = ctx.g
        dout = _rows(dout)
        dh = spmm(g, "csc", _lib.SPMM_WEIGHTED, dout, xg.dtype)
        dw = linear_wgrad(dh, xg)
        db = _take_colsum(dout)
        if db is None:
            db = colsum(dout).float()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = linear_dgrad(dh, wc)
            if dx.dtype != ctx.x_dtype:
                dx = cast(dx, ctx.x_dtype)
        if ctx.pad:
            dw = dw[:, :dw.size(1) - ctx.pad]
            dx = dx[:, :dx.size(1) - ctx.pad] if dx is not None else None
        return dx, dw, db, None, None, None


class GcnOutFn(torch.autograd.Function):
    """GCNConv with <= 4 output channels (the `hidden -> 2` logits layer of GCNNet, gnn.py:23), evaluated through the
    same narrow kernels as the SAGE logits layer: p = h W^T (one pass over h), out_i = sum_j rn(w_ji p_j) + b at
    width C, and the mirrored backward (csrc/sage_out.cu).  fp32 weights / accumulation / logits."""

    @staticmethod
    def forward(ctx, x, w, b, g: Graph):
        x = _rows(x)
        N, K = x.shape
        C = w.size(0)
        L = lib()
        # [W ; 0]: the zero rows absorb the `dout` half of dp in the shared skinny wgrad / dgrad kernels
        wcat = torch.cat([w.detach().float(), w.new_zeros((C, K), dtype=torch.float32)], dim=0).contiguous()
        p = torch.empty((N, C), dtype=torch.float32, device=x.device)
        check(L.egnn_skinny_project(ptr(x), dt(x), _ld(x), N, K, ptr(wcat), C, ptr(p), stream()))
        out = torch.empty((N, C), dtype=torch.float32, device=x.device)
        tmp = torch.empty((g.cap, C), dtype=torch.float32, device=x.device)
        bias = b.detach().float().contiguous() if b is not None else None
        check(L.egnn_gcn_out_fwd(ptr(g.csr_ptr), ptr(g.csr_src), ptr(g.w_csr), ptr(p), ptr(bias), C, ptr(out), N,
                                 ptr(tmp), g.cap, stream()))
        ctx.g, ctx.C = g, C
        ctx.save_for_backward(x, wcat)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, wcat = ctx.saved_tensors
        g, C = ctx.g, ctx.C
        N, K = x.shape
        L = lib()
        dout = _rows(dout).contiguous()
        dp = torch.empty((N, 2 * C), dtype=torch.float32, device=x.device)
        tmp = torch.empty((g.cap, C), dtype=torch.float32, device=x.device)
        check(L.egnn_gcn_out_bwd(ptr(g.csc_ptr), ptr(g.csc_dst), ptr(g.w_csc), ptr(dout), dt(dout), C, ptr(dp), N,
                                 ptr(tmp), g.cap, stream()))
        dw = torch.empty((2 * C, K), dtype=torch.float32, device=x.device)
        dsum = torch.empty(2 * C, dtype=torch.float32, device=x.device)
        ws = torch.empty(L.egnn_skinny_wgrad_workspace_floats(N, K, 2 * C), dtype=torch.float32, device=x.device)
        check(L.egnn_skinny_wgrad(ptr(x), dt(x), _ld(x), ptr(dp), 2 * C, N, K, ptr(dw), ptr(dsum), ptr(ws),
                                  stream()))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((N, K), dtype=x.dtype, device=x.device)
            check(L.egnn_skinny_dgrad(ptr(dp), ptr(wcat), 2 * C, ptr(dx), dt(dx), K, N, K, stream()))
        return dx, dw[:C], dsum[C:], None


# ---------------------------------------------------------------------------- GATConv ----
class GatConvFn(torch.autograd.Function):
    """PyG GATConv (heads, concat | head-mean) on the self-loop graph."""

    @staticmethod
    def forward(ctx, x, w, att_src, att_dst, bias, g: Graph, H: int, C: int, concat: bool, slope: float,
                bf16: bool):
        cd = torch.bfloat16 if bf16 else torch.float32
        x = _rows(x)
        N = x.size(0)
        pad = (-x.size(1)) % 8 if x.dtype == torch.float32 else 0   # 16-byte rows: vectorised / TMA operand loads
        if pad:
            x, w = pad_features(x, x.size(1) + pad, want_twin=bf16, twin_read_only=True), _pad_cols(w, pad)
        ctx.pad = pad
        xg = to_compute(x, cd)
        wc = w if w.dtype == cd else cast(w, cd)
        L = lib()
        K = xg.size(1)
        # the `hidden -> 2` logits layer (gnn.py:67): a [N, K] x [K, 2] product is a streaming pass over h, not a GEMM
        # -- the narrow kernels of the SAGE / GCN logits layers (csrc/sage_out.cu), fp32 accumulation
        ctx.skinny = H * C in (2, 4, 8) and K % 8 == 0 and K <= 1024 and _ld(xg) % 8 == 0
        if ctx.skinny:
            wc = (wc if wc.dtype == torch.float32 else cast(wc, torch.float32)).contiguous()  # bf16-rounded under autocast
            xs = torch.empty((N, H * C), dtype=torch.float32, device=x.device)
            check(L.egnn_skinny_project(ptr(xg), dt(xg), _ld(xg), N, K, ptr(wc), H * C, ptr(xs), stream()))
        else:
            xs = linear_fwd(xg, wc, out_dtype=torch.float32)  # [N, H*C]; attention path stays fp32
        dev = x.device
        a_s = torch.empty((N, H), dtype=torch.float32, device=dev)
        a_d = torch.empty((N, H), dtype=torch.float32, device=dev)
        asrc, adst = att_src.reshape(H * C).contiguous(), att_dst.reshape(H * C).contiguous()
        check(L.egnn_gat_scores(ptr(xs), N, H, C, ptr(asrc), ptr(adst), ptr(a_s), ptr(a_d), stream()))
        alpha = torch.empty((g.cap, H), dtype=torch.float32, device=dev)
        out = torch.empty((N, H * C if concat else C), dtype=torch.float32, device=dev)
        check(L.egnn_gat_fwd(ptr(g.csr_ptr), ptr(g.csr_src), ptr(xs), ptr(a_s), ptr(a_d), float(slope), H, C,
                             int(concat), ptr(bias), ptr(alpha), ptr(out), N, stream()))
        ctx.g, ctx.cfg, ctx.x_dtype = g, (H, C, concat, slope), x.dtype
        ctx.save_for_backward(xg, wc, xs, a_s, a_d, alpha, asrc, adst)
        return out

    @staticmethod
    def backward(ctx, dout):
        xg, wc, xs, a_s, a_d, alpha, asrc, adst = ctx.saved_tensors
        g = ctx.g
        H, C, concat, slope = ctx.cfg
        N, dev = xs.size(0), xs.device
        dout = _rows(dout)
        if dout.dtype != torch.float32:
            dout = cast(dout, torch.float32)
        dout = dout.contiguous()
        L = lib()
        f32 = dict(dtype=torch.float32, device=dev)
        dpre = torch.empty((g.cap, H), **f32)
        da_d = torch.empty((N, H), **f32)
        check(L.egnn_gat_bwd_dst(ptr(g.csr_ptr), ptr(g.csr_src), ptr(xs), ptr(a_s), ptr(a_d), ptr(alpha),
                                 ptr(dout), float(slope), H, C, int(concat), ptr(dpre), ptr(da_d), N, stream()))
        dxs = torch.empty((N, H * C), **f32)
        da_s = torch.empty((N, H), **f32)
        check(L.egnn_gat_bwd_src(ptr(g.csc_ptr), ptr(g.csc_dst), ptr(g.csc_pos), ptr(alpha), ptr(dpre),
                                 ptr(dout), ptr(da_d), ptr(asrc), ptr(adst), H, C, int(concat), ptr(dxs),
                                 ptr(da_s), N, stream()))
        datt = torch.empty((2, H * C), dtype=torch.float64, device=dev)
        ws = torch.empty(L.egnn_colreduce_workspace_bytes(H * C), dtype=torch.uint8, device=dev)
        check(L.egnn_gat_att_grad(ptr(xs), ptr(da_s), ptr(da_d), N, H, C, datt[0].data_ptr(),
                                  datt[1].data_ptr(), ptr(ws), stream()))
        datt = datt.float()
        # under bf16 autocast the Linear's backward runs in bf16 (PyG: grad of an autocast matmul): cast the fp32
        # attention-path gradient once so that wgrad / dgrad take the tensor-core kernels
        dbias = colsum(dout).float()
        dx = None
        if ctx.skinny:
            P, K = H * C, xg.size(1)
            dw = torch.empty((P, K), **f32)
            dsum = torch.empty(P, **f32)
            wsk = torch.empty(L.egnn_skinny_wgrad_workspace_floats(N, K, P), **f32)
            check(L.egnn_skinny_wgrad(ptr(xg), dt(xg), _ld(xg), ptr(dxs), P, N, K, ptr(dw), ptr(dsum), ptr(wsk), stream()))
            if ctx.needs_input_grad[0]:
                dx = torch.empty((N, K), dtype=xg.dtype, device=dev)
                check(L.egnn_skinny_dgrad(ptr(dxs), ptr(wc), P, ptr(dx), dt(dx), K, N, K, stream()))
        else:
            dxs_c = dxs if dxs.dtype == xg.dtype else cast(dxs, xg.dtype)
            dw = linear_wgrad(dxs_c, xg)
            if ctx.needs_input_grad[0]:
                dx = linear_dgrad(dxs_c, wc, out_dtype=xg.dtype)
        if dx is not None and dx.dtype != ctx.x_dtype:
            dx = cast(dx, ctx.x_dtype)
        if ctx.pad:
            dw = dw[:, :dw.size(1) - ctx.pad]
            dx = dx[:, :dx.size(1) - ctx.pad] if dx is not None else None
        return (dx, dw, datt[0].view(1, H, C), datt[1].view(1, H, C), dbias, None, None, None, None, None,
                None)


# ------------------------------------------------- BN + activation + dropout + residual ---
class DropoutState:
    """Per-model dropout stream: `seed` is fixed, `offset` (device int64) advances once per
    training forward so every step -- including CUDA-graph replays -- draws a fresh mask."""

    def __init__(self, seed: int, device):
        self.seed = int(seed) & 0x7FFFFFFFFFFFFFFF
        self.offset = torch.zeros(1, dtype=torch.int64, device=device)

    def advance(self, inc: int = 1):
        check(lib().egnn_counter_add(ptr(self.offset), int(inc), stream()))


class StatsReducer:
    """Hook for timestep-sharded runs: sums BatchNorm partial statistics over ranks.
    `n_total` is the global row count.  The default is the single-GPU identity."""

    def __init__(self, n_total: Optional[int] = None, group=None):
        self.n_total, self.group = n_total, group

    def reduce_(self, buf: torch.Tensor) -> torch.Tensor:
        if self.group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                      and self.n_total is not None):
            torch.distributed.all_reduce(buf, group=self.group)
        return buf


class BnActDropResFn(torch.autograd.Function):
    """y = dropout(act(batchnorm(z))) + res -- `src/models/gnn.py:186-192`.  bn is optional
    (gamma None), res is optional.  Training-mode BN uses batch statistics over all rows
    (globally reduced through `reducer` when sharded) and updates the running buffers."""

    @staticmethod
    def forward(ctx, z, res, gamma, beta, running_mean, running_var, training: bool, act: int, p: float,
                drop: Optional[DropoutState], layer: int, row0: int, eps: float, momentum: float,
                reducer: Optional[StatsReducer]):
        z = _rows(z)
        N, F = z.shape
        L = lib()
        dev = z.device
        res = _rows(res) if res is not None else None
        if res is not None and res.dtype != z.dtype:
            res = cast(res, z.dtype)
        strided_ok = _lean4(F, z, res)   # the streaming kernels take per-tensor row strides
        if not strided_ok:
            z = z.contiguous()
            res = res.contiguous() if res is not None else None
        use_bn = gamma is not None
        mean = rstd = None
        n_total = float(reducer.n_total) if (reducer is not None and reducer.n_total) else float(N)
        if use_bn:
            if training:
                st = colsum(z, want_sq=True)
                if reducer is not None:
                    reducer.reduce_(st)
                mean = torch.empty(F, dtype=torch.float32, device=dev)
                rstd = torch.empty(F, dtype=torch.float32, device=dev)
                check(L.egnn_bn_finalize(st[0].data_ptr(), st[1].data_ptr(), n_total, F, float(eps),
                                         float(momentum), ptr(mean), ptr(rstd), ptr(running_mean),
                                         ptr(running_var), stream()))
            else:
                mean = running_mean
                rstd = torch.rsqrt(running_var + eps)
        p_eff = float(p) if (training and p > 0) else 0.0
        if strided_ok and z.dtype == torch.bfloat16 and F % 8 == 0:
            _, y = alloc_cat(N, F, z.dtype, dev)   # [agg | y]: the next SAGEConv aggregates in place
        else:
            y = torch.empty((N, F), dtype=z.dtype, device=dev)
        seed = drop.seed if drop is not None else 0
        soff = drop.offset if drop is not None else None
        # per 4 columns one byte saved for the backward: low nibble = dropout keep bits (the backward skips Philox),
        # high nibble = ReLU gates (the backward skips the affine map)
        kb = (torch.empty((N, F // 4), dtype=torch.uint8, device=dev)
              if (strided_ok and training and z.requires_grad) else None)
        check(L.egnn_bn_act_dropout_res_fwd(ptr(z), ptr(res), ptr(y), dt(z), _ld(z), N, F, ptr(mean), ptr(rstd),
                                            ptr(gamma), ptr(beta), act, p_eff, seed, ptr(soff), layer, row0,
                                            _ld(res) if res is not None else 0, _ld(y), ptr(kb), None, None,
                                            stream()))
        ctx.kb = kb
        ctx.cfg = (use_bn, act, p_eff, seed, layer, row0, n_total, reducer, res is not None)
        ctx.soff = soff
        ctx.save_for_backward(z, mean, rstd, gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, mean, rstd, gamma, beta = ctx.saved_tensors
        use_bn, act, p_eff, seed, layer, row0, n_total, reducer, has_res = ctx.cfg
        soff = ctx.soff
        N, F = z.shape
        L = lib()
        dy = _rows(dy)
        if dy.dtype != z.dtype:
            dy = cast(dy, z.dtype)
        dy = dy.contiguous()
        dz = torch.empty((N, F), dtype=z.dtype, device=z.device)
        dgamma = dbeta = None
        if use_bn:
            sg = torch.empty((2, F), dtype=torch.float64, device=z.device)
            ws = torch.empty(L.egnn_colreduce_workspace_bytes(F), dtype=torch.uint8, device=z.device)
            check(L.egnn_bn_act_dropout_bwd_reduce(ptr(dy), ptr(z), dt(z), F, N, F, ptr(mean), ptr(rstd),
                                                   ptr(gamma), ptr(beta), act, p_eff, seed, ptr(soff), layer,
                                                   row0, sg[0].data_ptr(), sg[1].data_ptr(), ptr(ws), _ld(z),
                                                   ptr(ctx.kb), None, None, None, None, stream()))
            # d beta / d gamma: THIS rank's share (taken before the cross-rank sum the BatchNorm backward needs; the
            # weight-gradient all-reduce adds the ranks' shares -- returning the global sums would count them twice)
            sgf = sg.float()
            if reducer is not None:
                reducer.reduce_(sg)
            dzsum = torch.empty(F, dtype=torch.float32, device=z.device)
            ws2 = torch.empty(L.egnn_colreduce_workspace_bytes(F) + 64 * F, dtype=torch.uint8, device=z.device)
            check(L.egnn_bn_act_dropout_bwd_apply(ptr(dy), ptr(z), ptr(dz), dt(z), F, N, F, ptr(mean),
                                                  ptr(rstd), ptr(gamma), ptr(beta), act, p_eff, seed,
                                                  ptr(soff), layer, row0, sg[0].data_ptr(), sg[1].data_ptr(),
                                                  n_total, ptr(dzsum), ptr(ws2), _ld(z), ptr(ctx.kb), stream()))
            _publish_colsum(dz, dzsum)
            dbeta, dgamma = sgf[0], sgf[1]
        else:
            # the column sums of dz (the bias gradient of the conv that produced z) ride along, as in the BN branch
            dzsum = torch.empty(F, dtype=torch.float32, device=z.device)
            ws2 = torch.empty(L.egnn_colreduce_workspace_bytes(F) + 64 * F, dtype=torch.uint8, device=z.device)
            check(L.egnn_bn_act_dropout_bwd_apply(ptr(dy), ptr(z), ptr(dz), dt(z), F, N, F, None, None, None,
                                                  None, act, p_eff, seed, ptr(soff), layer, row0, None, None,
                                                  1.0, ptr(dzsum), ptr(ws2), _ld(z), ptr(ctx.kb), stream()))
            _publish_colsum(dz, dzsum)
        dres = dy if has_res else None
        return (dz, dres, dgamma, dbeta) + (None,) * 11


def act_dropout(z, act: int, p: float, training: bool, drop: Optional[DropoutState], layer: int,
                row0: int = 0):
    """dropout(act(z)) -- the ReLU/ELU + F.dropout pair of GCNNet/SAGENet/GATNet (gnn.py:29-30)."""
    return BnActDropResFn.apply(z, None, None, None, None, None, training, act, p, drop, layer, row0, 0.0,
                                0.0, None)


# ------------------------------------------------------------------------ time features ---
class InjectTimeFn(torch.autograd.Function):
    """[x | table[clamp(t-1)]] zero-padded to a multiple of 4 columns (gnn.py:168-179)."""

    @staticmethod
    def forward(ctx, x, t, table, width: int, want_twin: bool = False):
        x = _rows(x)
        if x.dtype != torch.float32:
            raise TypeError("node features must be float32")
        N, F = x.shape
        T, D = table.shape
        tb = table.detach().contiguous().float()
        out = torch.empty((N, width), dtype=torch.float32, device=x.device)
        # under bf16 autocast the same pass also writes the bf16 copy the GEMMs consume
        twin = None
        if want_twin:
            twin = (alloc_cat(N, width, torch.bfloat16, x.device)[1] if width % 8 == 0
                    else torch.empty((N, width), dtype=torch.bfloat16, device=x.device))
        check(lib().egnn_inject_time(ptr(x), _ld(x), ptr(t.contiguous()), ptr(tb), T, D, ptr(out), ptr(twin), width,
                                     _ld(twin) if twin is not None else 0, N, F, stream()))
        ctx.dims = (F, D, T)
        ctx.save_for_backward(t)
        if twin is None:
            twin = out.new_empty(0)
        ctx.mark_non_differentiable(twin)
        return out, twin

    @staticmethod
    def backward(ctx, dout, _dtwin=None):
        (t,) = ctx.saved_tensors
        F, D, T = ctx.dims
        dx = dout[:, :F] if ctx.needs_input_grad[0] else None
        dtab = None
        if ctx.needs_input_grad[2]:
            # learned nn.Embedding (not used by the BASELINE configs): deterministic segmented sum over the T rows
            dout = _rows(dout)
            L = lib()
            dtab = torch.empty((T, D), dtype=torch.float32, device=dout.device)
            ws = torch.empty(L.egnn_embed_grad_workspace_bytes(dout.size(0), T, D), dtype=torch.uint8,
                             device=dout.device)
            check(L.egnn_embed_grad(ptr(dout), dt(dout), _ld(dout), F, D, ptr(t.contiguous()), T, dout.size(0),
                                    ptr(dtab), ptr(ws), stream()))
        return dx, None, dtab, None, None


def inject_time(x, t, table, width: int):
    want_twin = amp_bf16()
    out, twin = InjectTimeFn.apply(x, t, table, width, want_twin)
    if want_twin:
        out._egnn_twin = twin
    return out


def append_scalar_time(x: torch.Tensor, timestep: torch.Tensor) -> torch.Tensor:
    """`use_time_scalar` (`src/train_gnn.py:314-317`): `cat([x, (timestep.float() / float(timestep.max())).unsqueeze(1)], 1)`
    on the device, bit-exact with the reference's CPU computation: the distinct quotients (one per timestep value) are
    computed on the host with the same torch division and gathered by `egnn_inject_time` as a [T, 1] table.
    One host read of (t_min, t_max), like the reference's `float(data.timestep.max())`."""
    if not (x.is_cuda and timestep.is_cuda):
        raise RuntimeError("egnn_b200 appends the time column on the GPU only (no CPU fallback)")
    x = _rows(x)
    N, F = x.shape
    t_min, t_max = (int(v) for v in torch.stack([timestep.min(), timestep.max()]).tolist())
    table = (torch.arange(t_min, t_max + 1, dtype=torch.int64).float() / float(t_max)).unsqueeze(1).to(x.device)
    t_rel = timestep if t_min == 1 else timestep - (t_min - 1)        # the kernel reads table[clamp(t - 1)]
    ld = (F + 1 + 7) // 8 * 8        # 16-byte rows (the kernel stores float4s); the view returned is [N, F + 1]
    out = torch.empty((N, ld), dtype=torch.float32, device=x.device)
    check(lib().egnn_inject_time(ptr(x), _ld(x), ptr(t_rel.contiguous()), ptr(table), table.size(0), 1, ptr(out), None,
                                 ld, 0, N, F, stream()))
    return out[:, :F + 1]


# ------------------------------------------------------------------------------- loss ----
TIME_SCHEMES = {"none": 0, "linear": 1, "sqrt": 2}


class MaskedCEFn(torch.autograd.Function):
    """`_make_loss_fn`'s per-row loss and mean (`src/train_gnn.py:136-176,201`) over precomputed train-row indices
    (`idx=None`: all rows, for logits that were masked already): class-weighted CE, or the focal variant
    (`focal_gamma >= 0`), times the optional time weight.  Forward also produces d loss / d logits, so the backward
    is a scale."""

    @staticmethod
    def forward(ctx, logits, y, idx, cw, n_total: float, focal_gamma: float = -1.0, timestep=None, t_min: float = 0.0,
                t_max: float = 1.0, time_scheme: int = 0):
        logits = _rows(logits).contiguous()
        if logits.size(1) != 2:
            raise ValueError("masked_weighted_ce is specialised for the reference's 2 classes")
        L = lib()
        N = logits.size(0)
        n_idx = N if idx is None else idx.numel()
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        dlog = torch.empty_like(logits)
        ws = torch.empty(L.egnn_ce_workspace_floats(n_idx), dtype=torch.float32, device=logits.device)
        check(L.egnn_masked_loss(ptr(logits), dt(logits), N, ptr(y), ptr(idx), n_idx, ptr(cw), float(n_total),
                                 float(focal_gamma), ptr(timestep), float(t_min), float(t_max), int(time_scheme),
                                 ptr(loss), ptr(dlog), ptr(ws), stream()))
        ctx.save_for_backward(dlog)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, gout):
        (dlog,) = ctx.saved_tensors
        return (dlog * gout.to(dlog.dtype),) + (None,) * 9


class L2MeanPenaltyFn(torch.autograd.Function):
    """loss + lambda * mean(w^2)  (`embed_l2 * model.time_emb.weight.pow(2).mean()`, `src/train_gnn.py:178-180`)."""

    @staticmethod
    def forward(ctx, loss, w, lam: float):
        out = loss.detach().clone().reshape(1)
        wc = w.detach().contiguous()
        check(lib().egnn_l2_mean_penalty(ptr(wc), wc.numel(), float(lam), ptr(out), None, stream()))
        ctx.save_for_backward(wc)
        ctx.lam = float(lam)
        return out.squeeze(0)

    @staticmethod
    def backward(ctx, gout):
        (wc,) = ctx.saved_tensors
        return gout, gout * (2.0 * ctx.lam / wc.numel()) * wc, None


def masked_weighted_ce(logits, y, train_idx, cw, n_total: Optional[float] = None):
    n_total = float(train_idx.numel()) if n_total is None else float(n_total)
    return MaskedCEFn.apply(logits, y, train_idx, cw, n_total)


def make_loss_fn(cfg: dict, cw: torch.Tensor, model, t_min, t_max):
    """Drop-in for `_make_loss_fn(cfg, cw, model, t_min, t_max)` (`src/train_gnn.py:136-183`): returns
    `loss_fn(logits, target, t_idx=None) -> scalar` over ALREADY-MASKED logits / targets, honouring `focal_loss`,
    `focal_gamma`, `time_loss_weighting` in {none, linear, sqrt} and `time_embed_l2`; one kernel for the per-row loss,
    its mean and d loss / d logits.  `loss_fn.on_rows(logits_all, y_all, idx, timestep_all, n_total)` is the same
    loss addressed through train-row indices (no boolean-mask gathers; what `TrainStep` calls)."""
    scheme = str(cfg.get("time_loss_weighting", "none"))
    if scheme not in TIME_SCHEMES:
        raise ValueError(f"unknown time_loss_weighting={scheme}")
    embed_l2 = float(cfg.get("time_embed_l2", 0.0))
    gamma = float(cfg.get("focal_gamma", 2.0)) if bool(cfg.get("focal_loss", False)) else -1.0

    def _l2(loss):
        te = getattr(model, "time_emb", None)
        if embed_l2 > 0.0 and te is not None:
            loss = L2MeanPenaltyFn.apply(loss, te.weight, embed_l2)
        return loss

    def loss_fn(logits, target, t_idx=None):
        sc = TIME_SCHEMES[scheme] if t_idx is not None else 0
        cwd = cw.to(logits.device)
        return _l2(MaskedCEFn.apply(logits, target, None, cwd, float(logits.size(0)), gamma, t_idx, float(t_min),
                                    float(t_max), sc))

    def on_rows(logits, y, idx, timestep=None, n_total=None):
        sc = TIME_SCHEMES[scheme] if timestep is not None else 0
        n = float(idx.numel()) if n_total is None else float(n_total)
        return _l2(MaskedCEFn.apply(logits, y, idx, cw.to(logits.device), n, gamma, timestep, float(t_min),
                                    float(t_max), sc))

    loss_fn.on_rows = on_rows
    loss_fn.cw = cw
    loss_fn.spec = dict(focal_gamma=gamma, time_scheme=TIME_SCHEMES[scheme], t_min=float(t_min), t_max=float(t_max),
                        embed_l2=embed_l2)
    return loss_fn
