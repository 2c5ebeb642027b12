import sys, os, torch
sys.path.insert(0, '/root/repo')
import egnn_b200 as E
from egnn_b200 import synthetic, ops, _lib
torch.cuda.set_device(0)
gr = synthetic.make_elliptic_like()
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
N = gr.num_nodes
def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
g = E.build_graph(ei.cuda(), N)
# an "identity-like" graph: every row gathers exactly itself (coalesced, no reuse)
self_ei = torch.arange(N).repeat(2, 1)
g_self = E.build_graph(self_ei.cuda(), N)
for F, di, do in [(64, torch.bfloat16, torch.bfloat16), (168, torch.float32, torch.bfloat16)]:
    xs = [torch.randn(N, F, device='cuda').to(di) for _ in range(8)]
    out = torch.empty(N, F, device='cuda', dtype=do)
    i = [0]
    def rot():
        i[0] += 1; ops.spmm(g, 'csr', _lib.SPMM_MEAN, xs[i[0] % 8], do, out=out)
    def same():
        ops.spmm(g, 'csr', _lib.SPMM_MEAN, xs[0], do, out=out)
    def selfg():
        i[0] += 1; ops.spmm(g_self, 'csr', _lib.SPMM_MEAN, xs[i[0] % 8], do, out=out)
    print(f"F={F}: rotating inputs {timeit(rot):.1f} us | same input (L2 warm) {timeit(same):.1f} us | self-loop graph (deg 1, sequential) {timeit(selfg):.1f} us")
# uniform-degree random graphs: separates "random access" from "degree variance"
t = gr.timestep
torch.manual_seed(0)
bounds = torch.cat([torch.zeros(1, dtype=torch.int64), torch.nonzero(t[1:] != t[:-1]).view(-1) + 1, torch.tensor([N])])
lo = bounds[torch.bucketize(torch.arange(N), bounds[1:], right=True)]
hi = bounds[torch.bucketize(torch.arange(N), bounds[1:], right=True) + 1]
for d in (1, 2, 4):
    src = (lo.repeat(d) + (torch.rand(N * d) * (hi - lo).repeat(d).float()).long()).clamp(max=N - 1)
    dst = torch.arange(N).repeat(d)
    gu = E.build_graph(torch.stack([src, dst]).cuda(), N)
    for F, di, do in [(64, torch.bfloat16, torch.bfloat16), (168, torch.float32, torch.bfloat16)]:
        xs = [torch.randn(N, F, device='cuda').to(di) for _ in range(8)]
        out = torch.empty(N, F, device='cuda', dtype=do)
        i = [0]
        def rot():
            i[0] += 1; ops.spmm(gu, 'csr', _lib.SPMM_MEAN, xs[i[0] % 8], do, out=out)
        print(f"uniform degree {d}, random in-timestep sources, F={F}: {timeit(rot):.1f} us")
