"""Sweep of the streaming SpMM kernel (csrc/spmm_stream.cu) against the per-row lean kernel.
Needs a library built with EXTRA=-DEGNN_SPMM_EXPERIMENT (environment re-read on every launch).
Prints time, compulsory-byte throughput and an output checksum (must be identical across kernels)."""
import os, sys, json, torch
sys.path.insert(0, '/root/repo')
import egnn_b200 as E
from egnn_b200 import synthetic, ops, _lib
torch.cuda.set_device(0)
rep = int(os.environ.get("PROBE_REPLICAS", "1"))
gr = synthetic.make_elliptic_like(train_window_k=8)
if rep > 1:
    gr = synthetic.replicate(gr, rep)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
N = gr.num_nodes
g = E.build_graph(ei, N)
PEAK = 6551e9


def timeit(fn, n=20, warm=3):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n): fn(i)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def case(F, dt_in, dt_out, view, nbuf):
    torch.manual_seed(F)
    xs = [torch.randn(N, F, device='cuda').to(dt_in) for _ in range(nbuf)]
    out = torch.empty(N, F, device='cuda', dtype=dt_out)
    mode = _lib.SPMM_MEAN if view == 'csr' else _lib.SPMM_SUM
    ms = timeit(lambda i: ops.spmm(g, view, mode, xs[i % nbuf], dt_out, out=out))
    ops.spmm(g, view, mode, xs[0], dt_out, out=out)
    chk = int(out.view(torch.int16 if dt_out == torch.bfloat16 else torch.int32).to(torch.int64).sum())
    by = N * F * (xs[0].element_size() + out.element_size()) + 4 * g.n_edges + 4 * (N + 1)
    return ms * 1e3, by / (ms * 1e-3) / PEAK, chk


cases = [("F168 f32->bf16 csr", 168, torch.float32, torch.bfloat16, 'csr', 3 if rep == 1 else 1),
         ("F64 bf16 csr", 64, torch.bfloat16, torch.bfloat16, 'csr', 8 if rep == 1 else 1),
         ("F64 bf16 csc", 64, torch.bfloat16, torch.bfloat16, 'csc', 8 if rep == 1 else 1)]
sweeps = [("lean", dict(EGNN_SPMM_IMPL="lean"))]
for W in ("2",):
    for cfg in ("16,3,4", "8,6,2", "8,6,3", "8,6,4", "4,11,2"):
        sweeps.append((f"stream168 W={W} G,VPL,D={cfg}", dict(EGNN_SPMM_IMPL="stream", EGNN_STREAM_W=W, EGNN_STREAM_CFG=cfg, only=0)))
for W in ("2",):
    for cfg in ("8,1,8", "4,2,4", "4,2,6", "4,2,8"):
        sweeps.append((f"stream64 W={W} G,VPL,D={cfg}", dict(EGNN_SPMM_IMPL="stream", EGNN_STREAM_W=W, EGNN_STREAM_CFG=cfg, only=1)))
if os.environ.get("PROBE_ONLY64"):      # wave / shape sweep of the narrow bf16 launches only (use with PROBE_REPLICAS=8)
    sweeps = [(f"stream64 W={W} G,VPL,D={cfg}", dict(EGNN_SPMM_IMPL="stream", EGNN_STREAM_W=W, EGNN_STREAM_CFG=cfg, only=1))
              for W in ("1", "2", "4") for cfg in ("8,1,8", "4,2,4", "4,2,6", "4,2,8")]
if os.environ.get("PROBE_WAVES"):       # waves sweep with the default shapes, all widths
    cases.insert(1, ("F128 bf16 csr", 128, torch.bfloat16, torch.bfloat16, 'csr', 4 if rep == 1 else 1))
    sweeps = [(f"stream W={W}", dict(EGNN_SPMM_IMPL="stream", EGNN_STREAM_W=W)) for W in os.environ["PROBE_WAVES"].split(",")]
if os.environ.get("PROBE_TASKS"):       # tasks-per-group sweep with the default shapes, all widths
    cases.insert(1, ("F128 bf16 csr", 128, torch.bfloat16, torch.bfloat16, 'csr', 4 if rep == 1 else 1))
    sweeps = [(f"stream T={T}", dict(EGNN_SPMM_IMPL="stream", EGNN_STREAM_T=T)) for T in os.environ["PROBE_TASKS"].split(",")]
sweeps.append(("stream default", dict(EGNN_SPMM_IMPL="stream")))
for name, env in sweeps:
    only = env.pop("only", None)
    for k in ("EGNN_SPMM_IMPL", "EGNN_STREAM_W", "EGNN_STREAM_CFG", "EGNN_STREAM_T"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for ci, (cname, F, di, do, view, nbuf) in enumerate(cases):
        if only == 0 and ci != 0: continue
        if only == 1 and ci == 0: continue
        us, frac, chk = case(F, di, do, view, nbuf)
        print(f"{name:34s} {cname:20s} {us:8.1f} us  frac {frac:.3f}  chk {chk}", flush=True)
