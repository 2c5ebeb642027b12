"""K6 tensor-core path: the tcgen05/TMA bf16 kernels against an fp64 matmul of the same
bf16-rounded operands (fp32 accumulation: only summation-order error remains) and against
the SIMT path; plus epilogue options (bias, accumulate, row-count division, fp32/bf16 out)."""
import pytest
import torch

from util import assert_close, rel_err

pytestmark = pytest.mark.gpu

TOL = 2e-6  # fp32 accumulation of exact bf16 products, relative to ||ref||_inf


def _mk(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).bfloat16().cuda()


@pytest.mark.parametrize("M", [1, 127, 128, 129, 5000, 203769])
@pytest.mark.parametrize("K,N", [(168, 64), (64, 64), (64, 168), (40, 32), (336, 128), (64, 8), (72, 24), (256, 256)])
def test_tn_fwd_matches_fp64(egnn, M, K, N):
    from egnn_b200 import ops
    if M == 203769 and (K, N) not in ((168, 64), (64, 168)):
        pytest.skip("full-size case kept to the bench shapes")
    a, w = _mk((M, K), 1), _mk((N, K), 2)
    ref = (a.double() @ w.double().t())
    out = ops.linear_fwd(a, w, out_dtype=torch.float32, impl=2)
    assert_close(out, ref, TOL, "tcgen05 fwd fp32 out")
    out_simt = ops.linear_fwd(a, w, out_dtype=torch.float32, impl=1)
    assert_close(out, out_simt, 2 * TOL, "tcgen05 vs SIMT")
    outb = ops.linear_fwd(a, w, out_dtype=torch.bfloat16, impl=2)
    assert torch.equal(outb, out.bfloat16()) or (outb.float() - ref.float()).abs().max() <= ref.abs().max() * 2 ** -8


def test_tn_epilogues(egnn):
    from egnn_b200 import ops
    M, K, N = 3000, 168, 64
    a, w = _mk((M, K), 3), _mk((N, K), 4)
    bias = torch.randn(N, device="cuda")
    ref = a.double() @ w.double().t() + bias.double()
    out = ops.linear_fwd(a, w, bias=bias, out_dtype=torch.float32, impl=2)
    assert_close(out, ref, TOL, "bias")
    base = torch.randn(M, N, device="cuda")
    acc = base.clone()
    ops.linear_fwd(a, w, out=acc, accumulate=True, impl=2)
    assert_close(acc, base.double() + a.double() @ w.double().t(), TOL, "accumulate fp32")
    accb = base.bfloat16()
    ops.linear_fwd(a, w, out=accb, accumulate=True, impl=2)
    assert_close(accb.float(), base.bfloat16().double() + a.double() @ w.double().t(), 2 ** -7, "accumulate bf16")
    # row-count division through a CSR pointer
    cnt = torch.randint(0, 5, (M,))
    ptr = torch.zeros(M + 1, dtype=torch.int32)
    ptr[1:] = torch.cumsum(cnt, 0)
    g, wt = _mk((M, N), 5), _mk((N, K), 6)
    d = ops.linear_dgrad(g, wt, out_dtype=torch.float32, row_div=ptr.cuda(), impl=2)
    refd = (g.double() @ wt.double()) / cnt.clamp(min=1).double().cuda().unsqueeze(1)
    assert_close(d, refd, TOL, "dgrad + row_div")
    d1 = ops.linear_dgrad(g, wt, out_dtype=torch.float32, row_div=ptr.cuda(), impl=1)
    assert_close(d1, refd, TOL, "SIMT dgrad + row_div")


@pytest.mark.parametrize("M", [1, 63, 64, 65, 6000, 203769])
@pytest.mark.parametrize("N,K", [(64, 168), (64, 64), (128, 128), (8, 64), (32, 168), (128, 256), (16, 40), (64, 336), (64, 384), (128, 336), (256, 256)])
def test_wgrad_matches_fp64(egnn, M, N, K):
    from egnn_b200 import ops
    if M == 203769 and (N, K) not in ((64, 168), (64, 64), (64, 336)):
        pytest.skip("full-size case kept to the bench shapes")
    g, x = _mk((M, N), 7), _mk((M, K), 8)
    ref = g.double().t() @ x.double()
    if M < 1024:
        out = ops.linear_wgrad(g, x)  # short reductions stay on the SIMT path
    else:
        out = ops.linear_wgrad(g, x, impl=2)
        again = ops.linear_wgrad(g, x, impl=2)
        assert torch.equal(out, again)  # deterministic partial reduction
    assert_close(out, ref, 5e-6, "wgrad")


def test_unsupported_layout_fails_loudly(egnn):
    from egnn_b200 import ops
    a, w = torch.randn(100, 30, device="cuda").bfloat16(), torch.randn(8, 30, device="cuda").bfloat16()
    with pytest.raises(RuntimeError, match="tcgen05"):
        ops.linear_fwd(a, w, impl=2)  # lda = 30 is not a multiple of 8 (TMA needs 16-byte strides)
    out = ops.linear_fwd(a, w)        # auto falls back to the SIMT *CUDA* kernel, never to CPU
    assert_close(out.float(), a.double() @ w.double().t(), 2 ** -7, "auto")


# ---- fp32 operands on the tensor cores: 3xTF32 ---------------------------------------------------
TOL_F32 = 5e-6   # rel ||ref||_inf vs fp64: a.w ~= a_lo.w_hi + a_hi.w_lo + a_hi.w_hi, fp32 accumulate (north_star bar: 1e-5)


@pytest.fixture(autouse=True)
def _f32_tensor_cores():
    """the fp32 sections below exercise the 3xTF32 kernels through the raw ops (the nets enable them for inference)"""
    from egnn_b200 import ops
    old, ops._F32_TC = ops._F32_TC, True
    yield
    ops._F32_TC = old


def _mk32(shape, seed, heavy=False):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(shape, generator=g)
    if heavy:   # a few columns with outliers, like the standardised Elliptic features
        t[:, ::17] *= torch.exp(0.75 * torch.randn((shape[0], t[:, ::17].shape[1]), generator=g))
    return t.cuda()


@pytest.mark.parametrize("M", [1024, 1025, 5000, 203769])
@pytest.mark.parametrize("K,N", [(168, 64), (64, 64), (336, 128), (128, 128), (64, 168), (128, 256), (40, 32), (12, 8)])
def test_fp32_tn_3xtf32_matches_fp64(egnn, M, K, N):
    from egnn_b200 import ops
    if M == 203769 and (K, N) not in ((168, 64), (336, 128)):
        pytest.skip("full-size case kept to the bench shapes")
    a, w = _mk32((M, K), 11, heavy=True), _mk32((N, K), 12) / K ** 0.5
    ref = a.double() @ w.double().t()
    out = ops.linear_fwd(a, w)                       # auto: 3xTF32 tcgen05 kernel for M >= 1024
    assert_close(out, ref, TOL_F32, "3xTF32 fwd")
    out_simt = ops.linear_fwd(a, w, impl=1)          # FFMA kernel
    assert_close(out_simt, ref, TOL_F32, "SIMT fwd")
    assert torch.equal(out, ops.linear_fwd(a, w))    # deterministic
    bias = torch.randn(N, device="cuda")
    outb = ops.linear_fwd(a, w, bias=bias)
    assert_close(outb, ref + bias.double(), TOL_F32, "3xTF32 fwd + bias")
    base = torch.randn(M, N, device="cuda")
    acc = base.clone()
    ops.linear_fwd(a, w, out=acc, accumulate=True)
    assert_close(acc, base.double() + ref, TOL_F32, "3xTF32 accumulate")


def test_fp32_dgrad_row_div_3xtf32(egnn):
    from egnn_b200 import ops
    M, K, N = 6000, 168, 64
    cnt = torch.randint(0, 5, (M,))
    ptr = torch.zeros(M + 1, dtype=torch.int32)
    ptr[1:] = torch.cumsum(cnt, 0)
    g, wt = _mk32((M, N), 13), _mk32((N, K), 14)
    d = ops.linear_dgrad(g, wt, row_div=ptr.cuda())
    refd = (g.double() @ wt.double()) / cnt.clamp(min=1).double().cuda().unsqueeze(1)
    assert_close(d, refd, TOL_F32, "3xTF32 dgrad + row_div")


@pytest.mark.parametrize("M", [1024, 1031, 6000, 203769])
@pytest.mark.parametrize("N,K", [(64, 168), (64, 64), (128, 128), (8, 64), (32, 168), (64, 336), (128, 336), (64, 384),
                                 (16, 40), (256, 128)])
def test_fp32_wgrad_3xtf32_matches_fp64(egnn, M, N, K):
    from egnn_b200 import ops
    if M == 203769 and (N, K) not in ((64, 168), (64, 336), (128, 128)):
        pytest.skip("full-size case kept to the bench shapes")
    g, x = _mk32((M, N), 21), _mk32((M, K), 22, heavy=True)
    ref = g.double().t() @ x.double()
    out = ops.linear_wgrad(g, x)                     # auto: 3xTF32 tcgen05 wgrad for M >= 1024
    assert_close(out, ref, TOL_F32, "3xTF32 wgrad")
    assert torch.equal(out, ops.linear_wgrad(g, x))  # deterministic partial reduction
    assert_close(ops.linear_wgrad(g, x, impl=1), ref, TOL_F32, "SIMT wgrad")


@pytest.mark.parametrize("K,N", [(336, 128), (168, 64), (128, 256), (64, 128), (40, 24)])
def test_fp32_exact_accumulation_modes(egnn, K, N):
    """`egnn_set_f32_tc_exact`: 1 = every 8-wide k-step product summed with IEEE adds outside the tensor core (what
    anything with a gradient runs; products wider than 128 columns as 128-column blocks), 0 = accumulate in TMEM (no-grad
    forwards).  Both inside the fp32 bar; the exact mode as close to fp64 as the FFMA kernel, and its error free of the
    truncation BIAS (mean signed error of positive sums ~0, where the in-TMEM mode leans towards zero)."""
    from egnn_b200 import _lib, ops
    L = _lib.lib()
    M = 20000
    a, w = _mk32((M, K), 31, heavy=True).abs(), _mk32((N, K), 32).abs() / K ** 0.5     # all-positive: sums never cancel
    ref = a.double() @ w.double().t()
    prev = L.egnn_set_f32_tc_exact(1)
    try:
        exact = ops.linear_fwd(a, w)
        assert L.egnn_set_f32_tc_exact(0) == 1
        fast = ops.linear_fwd(a, w)
        assert L.egnn_set_f32_tc_exact(1) == 0
        simt = ops.linear_fwd(a, w, impl=1)
        e_exact, e_fast, e_simt = (rel_err(t, ref) for t in (exact, fast, simt))
        assert e_exact <= TOL_F32 and e_fast <= TOL_F32
        assert e_exact <= max(3 * e_simt, 1.5e-6), (e_exact, e_simt)
        bias_exact = float(((exact.double() - ref) / ref).mean())
        bias_fast = float(((fast.double() - ref) / ref).mean())
        assert abs(bias_exact) < 2e-7, bias_exact
        if K >= 128:
            assert bias_fast < -3 * abs(bias_exact) - 1e-7, (bias_fast, bias_exact)   # truncation leans towards zero
        # epilogues through the exact path, incl. a 256-wide product split into column blocks
        cnt = torch.randint(0, 5, (M,))
        ptr = torch.zeros(M + 1, dtype=torch.int32)
        ptr[1:] = torch.cumsum(cnt, 0)
        half = N // 2 if N % 32 == 0 else 0
        d = torch.empty((M, N), device="cuda")
        ops._gemm(a, a.stride(0), 1, w, 1, w.stride(0), d, M, N, K, None, False, row_div=ptr.cuda(), row_div_cols=half)
        want = ref.clone()
        cols = slice(0, half) if half else slice(0, N)
        want[:, cols] = want[:, cols] / cnt.clamp(min=1).double().cuda().unsqueeze(1)
        assert_close(d, want, TOL_F32, "exact + row_div on the left columns")
        bias = torch.randn(N, device="cuda")
        base = torch.randn(M, N, device="cuda")
        acc = base.clone()
        ops.linear_fwd(a, w, bias=bias, out=acc, accumulate=True)
        assert_close(acc, base.double() + ref + bias.double(), TOL_F32, "exact + bias + accumulate")
    finally:
        L.egnn_set_f32_tc_exact(prev)


def test_tensor_core_gemm_from_a_fresh_thread(egnn):
    """The TMA descriptor encode is a driver call and needs a current context: a thread whose FIRST CUDA work is one of
    these GEMMs (the autograd worker when a backward starts with a weight gradient) must not fail with
    CUDA_ERROR_INVALID_CONTEXT."""
    import threading
    from egnn_b200 import ops
    g, x = _mk32((6000, 64), 41), _mk32((6000, 168), 42)
    a16, w16 = _mk((4096, 64), 43), _mk((64, 64), 44)
    torch.cuda.synchronize()
    res = {}

    def work(tag, fn):
        try:
            res[tag] = fn()
        except Exception as ex:          # noqa: BLE001 -- reported by the assert below
            res[tag] = ex
    old, ops._F32_TC = ops._F32_TC, True
    try:
        for tag, fn in (("f32 wgrad", lambda: ops.linear_wgrad(g, x)), ("f32 fwd", lambda: ops.linear_fwd(x, _mk32((64, 168), 45))),
                        ("bf16 fwd", lambda: ops.linear_fwd(a16, w16))):
            t = threading.Thread(target=work, args=(tag, fn))
            t.start()
            t.join()
            assert isinstance(res[tag], torch.Tensor), (tag, res[tag])
    finally:
        ops._F32_TC = old
    torch.cuda.synchronize()
    assert_close(res["f32 wgrad"], g.double().t() @ x.double(), TOL_F32, "wgrad from a fresh thread")
