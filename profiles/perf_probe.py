import sys, time, json, torch
sys.path.insert(0, '/root/repo')
import egnn_b200 as E
from egnn_b200 import synthetic, ops, _lib
from egnn_b200.train import TrainStep
torch.cuda.set_device(0)
gr = synthetic.make_elliptic_like(train_window_k=8)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
x = gr.x.cuda(); t = gr.timestep.cuda(); y = gr.y.cuda(); tm = gr.train_mask.cuda()
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
N = gr.num_nodes
t0 = time.time(); g = E.build_graph(ei, N); torch.cuda.synchronize(); print("graph build first", time.time() - t0)
print("graph build ms", timeit(lambda: E.build_graph(ei, N, validate=False)))
print("graph build +loops ms", timeit(lambda: E.build_graph(ei, N, self_loops=True, validate=False)))
for F, dt_in, dt_out in [(168, torch.float32, torch.float32), (168, torch.float32, torch.bfloat16), (64, torch.float32, torch.float32), (64, torch.bfloat16, torch.bfloat16), (128, torch.float32, torch.float32), (128, torch.bfloat16, torch.bfloat16)]:
    xx = torch.randn(N, F, device='cuda').to(dt_in)
    out = torch.empty(N, F, device='cuda', dtype=dt_out)
    ms = timeit(lambda: ops.spmm(g, 'csr', _lib.SPMM_MEAN, xx, dt_out, out=out))
    by = N * F * (xx.element_size() + out.element_size()) + 4 * g.n_edges + 4 * (N + 1)
    print(f"spmm mean F={F} {dt_in}->{dt_out}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s compulsory")
    ms = timeit(lambda: ops.spmm(g, 'csc', _lib.SPMM_DIV_NBR, xx, dt_out, out=out))
    print(f"spmm bwd  F={F}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s")
for (M, K, Nn, dtt) in [(N, 168, 64, torch.float32), (N, 168, 64, torch.bfloat16), (N, 64, 64, torch.float32), (N, 64, 2, torch.float32)]:
    a = torch.randn(M, K, device='cuda').to(dtt); w = torch.randn(Nn, K, device='cuda').to(dtt); gg = torch.randn(M, Nn, device='cuda').to(dtt)
    print(f"gemm fwd {M}x{K}x{Nn} {dtt}: {timeit(lambda: ops.linear_fwd(a, w))*1e3:.1f} us; dgrad {timeit(lambda: ops.linear_dgrad(gg, w))*1e3:.1f} us; wgrad {timeit(lambda: ops.linear_wgrad(gg, a))*1e3:.1f} us")
cfg = dict(hidden_dim=64, layers=3, dropout=0.2, time_embed_dim=2, time_embed_type='sin', max_timestep=49)
for amp in (False, True):
    torch.manual_seed(0)
    m = E.build_model('sage_resbn', 166, cfg).cuda()
    st = TrainStep(m, x, ei, t, y, tm, lr=5e-4, weight_decay=5e-5, amp=amp)
    n0 = _lib.launch_count(); st.run(); torch.cuda.synchronize(); print("launches/step", _lib.launch_count() - n0)
    print(f"rec_k8 amp={amp} eager step ms", timeit(st.run, n=10))
    st.capture()
    print(f"rec_k8 amp={amp} graphed step ms", timeit(st.run, n=20), "loss", float(st.loss))
