#!/usr/bin/env python
"""fp32 GEMM time per implementation (FFMA vs 3xTF32 tcgen05) at the layer shapes.  usage: python profiles/gemm_f32_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egnn_b200 as E
from egnn_b200 import ops

torch.cuda.set_device(0)
ops._F32_TC = True
N = 203769


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for (K, No) in ((168, 64), (336, 128), (128, 64), (168, 128), (128, 128), (64, 128)):
    As = [torch.randn(N, K, device="cuda") for _ in range(3)]
    W = torch.randn(No, K, device="cuda")
    it = [0]

    def run(impl):
        it[0] += 1
        return ops.linear_fwd(As[it[0] % 3], W, impl=impl)

    ref = As[0].double() @ W.double().t()
    e_tc = float((ops.linear_fwd(As[0], W).double() - ref).abs().max() / ref.abs().max())
    e_si = float((ops.linear_fwd(As[0], W, impl=1).double() - ref).abs().max() / ref.abs().max())
    byt = N * K * 4 + N * No * 4
    t_tc, t_si = timeit(lambda: run(None)), timeit(lambda: run(1))
    print(f"fwd [N,{K}]x[{No},{K}]: 3xTF32 {t_tc:.1f} us ({byt / t_tc / 1e3:.0f} GB/s, err {e_tc:.1e}) | FFMA {t_si:.1f} us (err {e_si:.1e})")
    G = torch.randn(N, No, device="cuda")
    t_w = timeit(lambda: ops.linear_wgrad(G, As[0]))
    t_ws = timeit(lambda: ops.linear_wgrad(G, As[0], impl=1), iters=5)
    refw = G.double().t() @ As[0].double()
    ew = float((ops.linear_wgrad(G, As[0]).double() - refw).abs().max() / refw.abs().max())
    ews = float((ops.linear_wgrad(G, As[0], impl=1).double() - refw).abs().max() / refw.abs().max())
    print(f"wgrad [{No},{K}] = G^T X: 3xTF32 {t_w:.1f} us (err {ew:.1e}) | FFMA {t_ws:.1f} us (err {ews:.1e})")
