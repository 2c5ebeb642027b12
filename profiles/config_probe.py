#!/usr/bin/env python
"""Step time of the other BASELINE configs (parity-test cases, not bench lines): CUDA-graph train step,
CUDA events, on the Elliptic-shaped graph or a k-times replicated one.
usage: python profiles/config_probe.py --arch sage_l3 --replicas 8 [--fp32] [--steps 20]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egnn_b200 as E
from egnn_b200 import synthetic
from egnn_b200.train import TrainStep

CFGS = {
    "gcn": dict(arch="gcn", hidden_dim=128, layers=3, dropout=0.5, lr=1e-3, weight_decay=5e-4, k=10, sym=False, ts=True),
    "sage": dict(arch="sage", hidden_dim=128, layers=2, dropout=0.5, lr=1e-3, weight_decay=5e-4, k=10, sym=True, ts=True),
    "gat": dict(arch="gat", hidden_dim=32, layers=2, heads=4, dropout=0.5, lr=1e-3, weight_decay=5e-4, k=10, sym=False, ts=True),
    "sage_l3": dict(arch="sage", hidden_dim=128, layers=3, dropout=0.4, lr=1e-3, weight_decay=5e-4, k=18, sym=True, ts=True),
    "rec_k8": dict(arch="sage_resbn", hidden_dim=64, layers=3, dropout=0.2, lr=5e-4, weight_decay=5e-5, k=8, sym=True,
                   ts=False, time_embed_dim=2, time_embed_type="sin", max_timestep=49),
}
ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="sage_l3")
ap.add_argument("--replicas", type=int, default=1)
ap.add_argument("--fp32", action="store_true")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--profile", action="store_true", help="per-kernel device time of eager steps (torch.profiler)")
args = ap.parse_args()
cfg = CFGS[args.arch]
torch.cuda.set_device(0)
t0 = time.time()
gr = synthetic.make_elliptic_like(train_window_k=cfg["k"])
if args.replicas > 1:
    gr = synthetic.replicate(gr, args.replicas)
x = gr.x
if cfg["ts"]:  # use_time_scalar: x || t / t.max()   (src/train_gnn.py:314-317)
    x = torch.cat([x, (gr.timestep.float() / gr.timestep.max().float()).unsqueeze(1)], dim=1)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1) if cfg["sym"] else gr.edge_index
print(f"# {args.arch} replicas={args.replicas} N={x.size(0)} F={x.size(1)} E={ei.size(1)} "
      f"{'fp32' if args.fp32 else 'bf16'} (host graph {time.time() - t0:.1f} s)", flush=True)
torch.manual_seed(42)
model = E.build_model(cfg["arch"], x.size(1), cfg).cuda()
model.set_dropout_seed(42, "cuda")
step = TrainStep(model, x.cuda(), ei.cuda(), gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(), lr=cfg["lr"],
                 weight_decay=cfg["weight_decay"], grad_clip=1.0, amp=not args.fp32)
del x, gr
step.run()
torch.cuda.synchronize()
if args.profile:
    import collections
    from torch.profiler import ProfilerActivity, profile
    step.run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            step.run()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in evs:
        agg[e.name[:110]][0] += 1
        agg[e.name[:110]][1] += e.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(evs) // 2} kernels/step, {tot / 2:.1f} us device time per step (eager)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{v[1] / 2:10.1f} us {v[0] // 2:4d}x {v[1] / tot * 100:5.1f}%  {k}")
    sys.exit(0)
step.capture(warmup=2)
for _ in range(3):
    step.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps):
    step.run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.steps
print(f"{args.arch} x{args.replicas}: {ms:.3f} ms/step, {ei.size(1) / ms / 1e6:.3f} GEdges/s, loss {float(step.loss):.4f}, "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
