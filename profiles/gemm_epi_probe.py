#!/usr/bin/env python
"""Time of the tcgen05 layer GEMM (egnn_linear_tc) per epilogue variant at the rec_k8 shapes (CUDA events, rotating
inputs larger than L2).  usage: python profiles/gemm_epi_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egnn_b200 as E
from egnn_b200 import _lib, fused
from egnn_b200._lib import lib, ptr

torch.cuda.set_device(0)
dev = torch.device("cuda")
N = 203769
L = lib()


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for (K, No, tag) in ((336, 128, "layer0 [N,336]x[128,336]"), (128, 64, "layer1 [N,128]x[64,128]"),
                     (64, 128, "dgrad  [N,64]x[128,64]")):
    As = [torch.randn(N, K, device=dev).bfloat16() for _ in range(4)]
    W = torch.randn(No, K, device=dev).bfloat16()
    out = torch.empty(N, No, dtype=torch.bfloat16, device=dev)
    bias = torch.randn(No, device=dev)
    n_parts = int(L.egnn_linear_stats_parts(N))
    parts = torch.empty(n_parts, 2, 64, device=dev)
    add = torch.randn(N, 64, device=dev).bfloat16()
    rowptr = torch.arange(N + 1, dtype=torch.int32, device=dev) * 2
    it = [0]

    def run(**kw):
        it[0] += 1
        fused._linear_tc(As[it[0] % 4], W, out, **kw)

    print(f"{tag}: plain {timeit(lambda: run()):.1f} us, +bias {timeit(lambda: run(bias=bias)):.1f} us, "
          f"+bias+stats64 {timeit(lambda: run(bias=bias, stats=parts, stats_cols=64)):.1f} us, "
          f"stats64 {timeit(lambda: run(stats=parts, stats_cols=64)):.1f} us"
          + (f", row_div {timeit(lambda: run(row_div=rowptr, row_div_cols=64)):.1f} us, row_div+addend "
             f"{timeit(lambda: run(row_div=rowptr, row_div_cols=64, addend=add, add_col0=64)):.1f} us" if No == 128 else ""))
