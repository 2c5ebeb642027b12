"""K6 tensor-core path: the tcgen05/TMA bf16 kernels against an fp64 matmul of the same
bf16-rounded operands (fp32 accumulation: only summation-order error remains) and against
the SIMT path; plus epilogue options (bias, accumulate, row-count division, fp32/bf16 out)."""
import pytest
import torch

from util import assert_close

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
