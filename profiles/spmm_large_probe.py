#!/usr/bin/env python
"""Mean-aggregation kernels on k-times replicated graphs (working set far beyond the 126 MB L2: the
HBM-honest roofline measurement).  usage: python profiles/spmm_large_probe.py 8 64"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egnn_b200 as E
from egnn_b200 import _lib, ops, synthetic

torch.cuda.set_device(0)
PEAK = 6551.0
base = synthetic.make_elliptic_like()
ei1 = torch.cat([base.edge_index, base.edge_index.flip(0)], 1)
n1 = base.num_nodes


def timeit(fn, n=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


for k in [int(v) for v in sys.argv[1:]] or [8]:
    ei = torch.cat([ei1 + r * n1 for r in range(k)], 1).cuda()
    N, Ee = n1 * k, ei.size(1)
    g = E.build_graph(ei, N)
    for F, di, do in [(168, torch.float32, torch.bfloat16), (128, torch.bfloat16, torch.bfloat16),
                      (64, torch.bfloat16, torch.bfloat16)]:
        x = torch.randn(N, F, device="cuda").to(di)
        out = torch.empty(N, F, device="cuda", dtype=do)
        t = timeit(lambda: ops.spmm(g, "csr", _lib.SPMM_MEAN, x, do, out=out))
        nbytes = N * F * (x.element_size() + out.element_size()) + 4 * Ee + 4 * (N + 1)
        print(f"x{k}: mean SpMM F={F} {str(di)[6:]}->{str(do)[6:]}: {t * 1e6:9.1f} us  algorithmic {nbytes / 1e6:8.1f} MB  "
              f"{nbytes / t / 1e9:7.1f} GB/s = {nbytes / t / 1e9 / PEAK:.3f} of measured HBM peak", flush=True)
        del x, out
    del g, ei
    torch.cuda.empty_cache()
