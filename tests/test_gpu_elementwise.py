"""K7 / step-tail parity: Philox mask (bit-exact vs NumPy), column reductions, BN+act+dropout+res
forward/backward, masked CE, clip+Adam, time injection."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import graph_build_np as G
from oracle import pyg_restated as O
from util import REL_FP32, assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", [0.2, 0.5])
@pytest.mark.parametrize("shape", [(1000, 64), (257, 130), (33, 2)])
def test_dropout_mask_bitexact(egnn, p, shape):
    from egnn_b200 import ops
    n, f = shape
    seed, layer, row0 = 0x1234_5678_9ABC, 3, (1 << 33) + 17
    m = ops.dropout_mask(n, f, p, seed, layer, row0).cpu().numpy()
    ref = G.dropout_keep_mask(seed, layer, row0, n, f, p)
    assert np.array_equal(m, ref)
    assert abs(m.mean() - (1 - p)) < 0.02 or n * f < 1000
    # shard invariance: rows [a, b) of the global mask == mask generated with row0 + a
    sub = ops.dropout_mask(100, f, p, seed, layer, row0 + 50).cpu().numpy() if n >= 150 else None
    if sub is not None:
        assert np.array_equal(sub, ref[50:150])
    off = torch.tensor([5], dtype=torch.int64, device="cuda")
    m2 = ops.dropout_mask(n, f, p, seed, layer, row0, seed_off=off).cpu().numpy()
    assert np.array_equal(m2, G.dropout_keep_mask(seed + 5, layer, row0, n, f, p))


@pytest.mark.parametrize("shape", [(5000, 64), (70001, 168), (1234, 2), (3, 7)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum(egnn, shape, dtype):
    from egnn_b200 import ops
    a = (torch.randn(shape) * 3 + 0.5).to(dtype)
    got = ops.colsum(a.cuda(), want_sq=True).cpu()
    ref = torch.stack([a.double().sum(0), (a.float() * a.float()).double().sum(0)])
    assert_close(got, ref, 2e-6, "colsum")
    again = ops.colsum(a.cuda(), want_sq=True).cpu()
    assert torch.equal(got, again)  # deterministic


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("use_bn,has_res,act,p", [(True, True, 1, 0.2), (True, False, 1, 0.0),
                                                  (False, False, 1, 0.5), (False, False, 2, 0.5),
                                                  (True, True, 2, 0.3)])
def test_bn_act_dropout_res(egnn, dtype, use_bn, has_res, act, p):
    from egnn_b200 import ops
    n, f = 4097, 64
    torch.manual_seed(0)
    z = (torch.randn(n, f) * 2 + 0.3).to(dtype)
    res = torch.randn(n, f).to(dtype) if has_res else None
    bn = torch.nn.BatchNorm1d(f)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    bn_ref = torch.nn.BatchNorm1d(f)
    bn_ref.load_state_dict(bn.state_dict())
    bn = bn.cuda()
    drop = ops.DropoutState(99, torch.device("cuda"))
    drop.advance(3)
    zc = z.cuda().requires_grad_(True)
    rc = res.cuda().requires_grad_(True) if has_res else None
    args = (bn.weight, bn.bias, bn.running_mean, bn.running_var) if use_bn else (None, None, None, None)
    y = ops.BnActDropResFn.apply(zc, rc, *args, True, act, p, drop, 1, 7, bn.eps, bn.momentum, None)
    gy = torch.randn(n, f).to(dtype)
    y.backward(gy.cuda())
    # oracle in fp32 on the same (possibly bf16-rounded) inputs with the very same mask
    mask = torch.from_numpy(G.dropout_keep_mask(99 + 3, 1, 7, n, f, p)).float() if p > 0 else None
    zr = z.float().requires_grad_(True)
    rr = res.float().requires_grad_(True) if has_res else None
    u = bn_ref(zr) if use_bn else zr
    a = F.relu(u) if act == 1 else F.elu(u)
    if mask is not None:
        a = a * mask * (1.0 / (1.0 - p))
    yr = a + rr if has_res else a
    yr.backward(gy.float())
    tol = REL_FP32 if dtype == torch.float32 else 1e-2
    assert_close(y.float(), yr, tol, "y")
    assert_close(zc.grad.float(), zr.grad, tol, "dz")
    if has_res:
        assert_close(rc.grad.float(), rr.grad, tol, "dres")
    if use_bn:
        assert_close(bn.weight.grad, bn_ref.weight.grad, 1e-3 if dtype == torch.bfloat16 else REL_FP32, "dgamma")
        assert_close(bn.bias.grad, bn_ref.bias.grad, 1e-3 if dtype == torch.bfloat16 else REL_FP32, "dbeta")
        assert_close(bn.running_mean, bn_ref.running_mean, REL_FP32, "running_mean")
        assert_close(bn.running_var, bn_ref.running_var, REL_FP32, "running_var")


def test_masked_ce(egnn):
    from egnn_b200 import ops
    n = 5000
    torch.manual_seed(1)
    logits = (torch.randn(n, 2) * 3).requires_grad_(True)
    y = torch.randint(-1, 2, (n,))
    mask = y >= 0
    cw = torch.tensor([0.6, 7.5])
    ref = O.masked_weighted_ce(logits, y, mask, cw)
    ref.backward()
    lc = logits.detach().cuda().requires_grad_(True)
    idx = torch.nonzero(mask).view(-1).cuda()
    got = ops.masked_weighted_ce(lc, y.cuda(), idx, cw.cuda())
    (got * 1.0).backward()
    assert_close(got, ref, REL_FP32, "loss")
    assert_close(lc.grad, logits.grad, REL_FP32, "dlogits")


def test_clip_adam_matches_torch(egnn):
    from egnn_b200.train import FlatClipAdam
    torch.manual_seed(2)
    shapes = [(64, 168), (64,), (2, 64), (7,)]
    ps_ref = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    ps = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ps_ref]
    opt_ref = torch.optim.Adam(ps_ref, lr=3e-3, weight_decay=1e-4)
    opt = FlatClipAdam(ps, lr=3e-3, weight_decay=1e-4, max_norm=1.0)
    for step in range(5):
        grads = [torch.randn(s) * (10.0 if step % 2 == 0 else 0.01) for s in shapes]
        opt.zero_grad()
        for p, pr, g in zip(ps, ps_ref, grads):
            pr.grad = g.clone()
            p.grad = g.cuda()
        opt.gather_grads()   # one multi-tensor copy into the flat gradient buffer
        total = torch.nn.utils.clip_grad_norm_(ps_ref, 1.0)
        opt_ref.step()
        opt.step()
        assert_close(opt.grad_norm.cpu()[0], total, REL_FP32, "grad norm")
        for p, pr in zip(ps, ps_ref):
            assert_close(p.data, pr.data, REL_FP32, f"param step {step}")


def test_inject_time_bitexact(egnn):
    from egnn_b200 import ops
    from egnn_b200.models import sinusoid_table
    n, f, T = 3000, 166, 49
    torch.manual_seed(3)
    x = torch.randn(n, f)
    t = torch.randint(-2, 60, (n,))  # includes out-of-range timesteps (clamped)
    net = O.SAGEResBNNet(f, 8, 2, time_embed_dim=2, time_embed_type="sin", max_timestep=T)
    ref = net._inject_time(x, t)
    table = sinusoid_table(T, 2)
    assert torch.equal(table, O.sinusoid_table(T, 2))
    got = ops.InjectTimeFn.apply(x.cuda(), t.cuda(), table.cuda(), 168)[0].cpu()
    assert torch.equal(got, ref)
    # scalar-time append of train_gnn.py:315-317 through the same kernel (F -> F+1, padded to 168)
    tt = torch.randint(1, 50, (n,))
    tn = (tt.float() / float(tt.max())).unsqueeze(1)
    tab = (torch.arange(1, T + 1).float() / float(tt.max())).unsqueeze(1)
    got = ops.InjectTimeFn.apply(x.cuda(), tt.cuda(), tab.cuda(), 168)[0].cpu()
    assert torch.equal(got[:, :167], torch.cat([x, tn], dim=1)) and (got[:, 167] == 0).all()
    # the public helper (builds the [T, 1] table itself), also for timestep ranges that do not start at 1
    for lo, hi in ((1, 50), (0, 37), (5, 12)):
        tt = torch.randint(lo, hi, (n,))
        ref = torch.cat([x, (tt.float() / float(tt.max())).unsqueeze(1)], dim=1)      # src/train_gnn.py:315-317 on the CPU
        got = egnn.append_scalar_time(x.cuda(), tt.cuda())
        assert got.shape == (n, f + 1) and torch.equal(got.cpu(), ref)


def test_p2p_allreduce_single_rank_protocol(egnn):
    """csrc/p2p.cu with world_size 1 (the only size a 1-GPU box can run): the rank pushes into its own buffer,
    flags itself and sums one slot -- exercises the epoch / parity / chunk-flag protocol over several calls,
    for fp64 statistics and a multi-chunk fp32 gradient vector.  The 2/4/8-rank behaviour is checked by
    profiles/shard_check.py on a multi-GPU box."""
    from egnn_b200 import _lib
    L = _lib.lib()
    for dtype, code, n_max, n in ((torch.float64, _lib.F64, 1024, 128), (torch.float32, _lib.F32, 41090, 41090)):
        nbytes = L.egnn_p2p_allreduce_buffer_bytes(1, n_max, code)
        buf = torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device="cuda")
        ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device="cuda")
        epoch = torch.zeros(2, dtype=torch.int64, device="cuda")    # {calls completed, ticket}
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        for it in range(5):
            x = torch.randn(n, dtype=dtype, device="cuda")
            want = x.clone()
            _lib.check(L.egnn_p2p_allreduce(x.data_ptr(), x.data_ptr(), n, code, n_max, ptrs.data_ptr(), 0, 1,
                                            epoch.data_ptr(), err.data_ptr(), 0, _lib.stream()))
            assert torch.equal(x, want)
        assert epoch.tolist() == [5, 0] and int(err.item()) == 0


def test_p2p_allreduce_missing_peer_fails_loudly(egnn):
    """a peer that never arrives must not yield a silently un-reduced vector.  World of 2 simulated on one
    GPU (two buffers, rank 1 never runs): after the (shortened) timeout the error flag is set and stays set, the
    output is NaN, and P2PAllReduce.check()-style polling of the flag raises."""
    from egnn_b200 import _lib
    L = _lib.lib()
    n_max = n = 128
    nbytes = L.egnn_p2p_allreduce_buffer_bytes(2, n_max, _lib.F64)
    bufs = [torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device="cuda") for _ in range(2)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    epoch = torch.zeros(2, dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    _lib.check(L.egnn_p2p_allreduce(x.data_ptr(), x.data_ptr(), n, _lib.F64, n_max, ptrs.data_ptr(), 0, 2,
                                    epoch.data_ptr(), err.data_ptr(), 20, _lib.stream()))
    torch.cuda.synchronize()
    assert int(err.item()) == 1
    assert torch.isnan(x).all()


def test_learned_time_embedding_backward_is_deterministic(egnn):
    """`time_embed_type: learned` (src/models/gnn.py:152,172-176): the table gradient is a fixed-order segmented
    sum (egnn_embed_grad), equal to the oracle's index_add within fp32 and bitwise reproducible."""
    from egnn_b200 import ops
    n, f, T, D = 20000, 166, 49, 6
    torch.manual_seed(4)
    x = torch.randn(n, f)
    t = torch.randint(-1, 55, (n,))
    table = torch.randn(T, D)
    w = torch.randn(n, 176)
    ref_tab = table.clone().requires_grad_(True)
    idx = torch.clamp(t - 1, 0, T - 1)
    ref = torch.cat([x, ref_tab[idx]], dim=1)
    (ref * w[:, :f + D]).sum().backward()
    grads = []
    for _ in range(2):
        tab = table.cuda().requires_grad_(True)
        out, _ = ops.InjectTimeFn.apply(x.cuda(), t.cuda(), tab, 176, False)
        assert torch.equal(out[:, :f + D].cpu(), ref.detach()) and (out[:, f + D:] == 0).all()
        (out * w.cuda()).sum().backward()
        grads.append(tab.grad.clone())
    assert torch.equal(grads[0], grads[1])
    assert_close(grads[0], ref_tab.grad, REL_FP32, "d time_emb.weight")
