"""Multi-GPU correctness: the timestep-sharded step -- peer-memory all-reduce of the gradients, the
single-kernel BatchNorm statistics exchanges, CUDA-graph replays -- against the single-GPU step on the same graph.
Needs two visible GPUs (the driver's one-GPU `-m gpu` run skips it; `gpurun --gpus 2 -- python -m pytest tests -m gpu
-k multi` runs it and profiles/r02/shard_check_2gpu.txt is its committed output)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_sharded_step_equals_single_gpu():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "profiles", "shard_check.py")], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout and r.stdout.count("-> OK") == 3


def test_fused_bn_exchange_kernels_single_rank(egnn):
    """egnn_bn_stats_exchange / egnn_bn_bwd_sums_exchange with world = 1 (what a one-GPU box can run): the reduce +
    exchange + finalise kernel equals the separate reduce / finalise kernels bit for bit, over several epochs."""
    from egnn_b200 import _lib
    L = _lib.lib()
    F, n_parts, N = 64, 148, 203769.0
    nbytes = L.egnn_p2p_allreduce_buffer_bytes(1, 1024, _lib.F64)
    buf = torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device="cuda")
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device="cuda")
    epoch = torch.zeros(2, dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    f32 = dict(dtype=torch.float32, device="cuda")
    for it in range(3):
        parts = torch.randn(n_parts, 2, F, **f32).abs() * 1000
        outs = []
        for fused in (False, True):
            mean, rstd = torch.empty(F, **f32), torch.empty(F, **f32)
            rm, rv = torch.full((F,), 0.25, **f32), torch.full((F,), 2.0, **f32)
            nb = torch.zeros(1, dtype=torch.int64, device="cuda")
            if fused:
                _lib.check(L.egnn_bn_stats_exchange(parts.data_ptr(), n_parts, F, N, 1e-5, 0.1, mean.data_ptr(),
                                                    rstd.data_ptr(), rm.data_ptr(), rv.data_ptr(), nb.data_ptr(), 1024,
                                                    ptrs.data_ptr(), 0, 1, epoch.data_ptr(), err.data_ptr(), 0,
                                                    _lib.stream()))
            else:
                _lib.check(L.egnn_bn_finalize_parts(parts.data_ptr(), n_parts, F, N, 1e-5, 0.1, mean.data_ptr(),
                                                    rstd.data_ptr(), rm.data_ptr(), rv.data_ptr(), nb.data_ptr(),
                                                    _lib.stream()))
            outs.append((mean, rstd, rm, rv, nb))
        for a, b in zip(*outs):
            assert torch.allclose(a.double(), b.double(), rtol=1e-6, atol=0), "fused forward exchange differs"
        assert int(outs[1][4]) == 1
        partial = torch.randn(296, 2, F, dtype=torch.float64, device="cuda")
        sums = torch.empty(2, F, dtype=torch.float64, device="cuda")
        g0, g1 = torch.empty(F, **f32), torch.empty(F, **f32)
        _lib.check(L.egnn_bn_bwd_sums_exchange(partial.data_ptr(), 296, F, sums.data_ptr(), g0.data_ptr(), g1.data_ptr(),
                                               1024, ptrs.data_ptr(), 0, 1, epoch.data_ptr(), err.data_ptr(), 0,
                                               _lib.stream()))
        want = partial.sum(0)
        assert torch.allclose(sums, want, rtol=1e-12, atol=1e-12)
        assert torch.equal(g0, sums[0].float()) and torch.equal(g1, sums[1].float())
    assert epoch.tolist() == [6, 0] and int(err.item()) == 0
