import sys, torch
sys.path.insert(0, '/root/repo')
from egnn_b200 import synthetic
torch.cuda.set_device(0)
gr = synthetic.make_elliptic_like()
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
N = gr.num_nodes
def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
for F, dt in [(64, torch.bfloat16), (168, torch.float32)]:
    xs = [torch.randn(N, F, device='cuda').to(dt) for _ in range(4)]
    out = torch.empty(N, F, device='cuda', dtype=dt)
    i = [0]
    def cp():
        i[0] += 1; out.copy_(xs[i[0] % 4])
    first_src = torch.zeros(N, dtype=torch.int64, device='cuda')
    first_src.scatter_(0, ei[1], ei[0])          # one in-neighbour per row (arbitrary)
    def g1():
        i[0] += 1; torch.index_select(xs[i[0] % 4], 0, first_src, out=out)
    oute = torch.empty(ei.size(1), F, device='cuda', dtype=dt)
    def gE():
        i[0] += 1; torch.index_select(xs[i[0] % 4], 0, ei[0], out=oute)
    print(f"F={F} {dt}: copy {timeit(cp):.1f} us | index_select N rows {timeit(g1):.1f} us | index_select E rows {timeit(gE):.1f} us")
