#!/usr/bin/env python
"""Per-kernel device time of one rec_k8 train step (torch.profiler / CUPTI; warm caches, unlike the
serialised ncu launch list).  usage: python profiles/step_breakdown.py [--fp32] [--steps 5] [--seq]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import egnn_b200 as E
from egnn_b200.train import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--fp32", action="store_true")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--seq", action="store_true", help="print the kernel sequence of the last step")
ap.add_argument("--arch", default="rec_k8")
ap.add_argument("--eval", action="store_true", help="profile the eval_split forward (fp32, no autocast) instead")
args = ap.parse_args()
torch.cuda.set_device(0)
gr = bench.host_graph(1)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
torch.manual_seed(42)
CFG = dict(bench.CFG)
model = E.build_model(CFG["arch"], gr.x.size(1), CFG).cuda()
model.set_dropout_seed(42, "cuda")
step = TrainStep(model, gr.x.cuda(), ei, gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(), lr=CFG["lr"],
                 weight_decay=CFG["weight_decay"], grad_clip=1.0, amp=not args.fp32)
if args.eval:
    from egnn_b200.train import eval_probs
    xd, td = gr.x.cuda(), gr.timestep.cuda()
    run = lambda: eval_probs(model, xd, ei, td)
else:
    run = step.run
for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps):
        run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    agg[e.name[:100]][0] += 1
    agg[e.name[:100]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"# {len(evs) // args.steps} kernels/step, {tot / args.steps:.1f} us device time per step (eager, warm)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / args.steps:9.1f} us {v[0] // args.steps:4d}x {v[1] / tot * 100:5.1f}%  {k}")
if args.seq:
    n = len(evs) // args.steps
    for e in sorted(evs, key=lambda e: e.time_range.start)[-n:]:
        print(f"   {e.device_time:8.1f}  {e.name[:110]}")
