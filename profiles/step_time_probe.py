import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import bench
import egnn_b200 as E
from egnn_b200 import fused
from egnn_b200.train import TrainStep
torch.cuda.set_device(0)
gr = bench.host_graph(1)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
torch.manual_seed(42)
CFG = dict(bench.CFG)
model = E.build_model(CFG["arch"], gr.x.size(1), CFG).cuda()
model.set_dropout_seed(42, "cuda")
step = TrainStep(model, gr.x.cuda(), ei, gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(), lr=CFG["lr"],
                 weight_decay=CFG["weight_decay"], grad_clip=1.0, amp=True)
def timed(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("eager ms", timed(step.run), "cache", fused.STATIC_INPUTS.hits, fused.STATIC_INPUTS.misses)
step.capture(warmup=2)
print("graph ms", timed(step.run), "cache", fused.STATIC_INPUTS.hits, fused.STATIC_INPUTS.misses)
step.capture_dynamic()
print("graph dynamic ms", timed(lambda: step.run(dynamic=True)))
# --- the same with bench.py's sharded context (world = 1) attached
from egnn_b200.shard import ShardedContext, make_shard
fused.STATIC_INPUTS.clear()
sh = make_shard(gr, 0, 1)
ctx = ShardedContext(sh, torch.device("cuda", 0))
torch.manual_seed(42)
model2 = ctx.attach(E.build_model(CFG["arch"], gr.x.size(1), CFG).cuda())
model2.set_dropout_seed(42, "cuda")
step2 = TrainStep(model2, gr.x.cuda(), ei, gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(), lr=CFG["lr"],
                  weight_decay=CFG["weight_decay"], grad_clip=1.0, amp=True, cw=ctx.class_weight,
                  n_train_total=ctx.n_train_total, health_check=ctx.check)
step2.run(); step2.run()
step2.capture(warmup=4)
print("graph + ctx ms", timed(step2.run), "cache", fused.STATIC_INPUTS.hits, fused.STATIC_INPUTS.misses)
from torch.profiler import ProfilerActivity, profile
import collections
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step2.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    agg[e.name[:90]][0] += 1; agg[e.name[:90]][1] += e.device_time
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{v[1] / 3:9.1f} us {v[0] // 3:4d}x  {k}")
