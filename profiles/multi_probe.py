#!/usr/bin/env python
"""Per-kernel time of the sharded rec_k8 step on rank 0 (weak scaling, one replica per rank), CUDA-graph replays.
launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/multi_probe.py"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench
import egnn_b200 as E
from egnn_b200.shard import ShardedContext, make_shard
from egnn_b200.train import TrainStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
CFG = dict(bench.CFG)
gr = bench.host_graph(world)
sh = make_shard(gr, rank, world)
lg = sh.graph
ctx = ShardedContext(sh, dev)
torch.manual_seed(42)
model = ctx.attach(E.build_model(CFG["arch"], lg.x.size(1), CFG).to(dev))
model.set_dropout_seed(42, dev)
for p in model.parameters():
    dist.broadcast(p.data, 0)
ei = torch.cat([lg.edge_index, lg.edge_index.flip(0)], 1).to(dev)
step = TrainStep(model, lg.x.to(dev), ei, lg.timestep.to(dev), lg.y.to(dev), lg.train_mask.to(dev), lr=CFG["lr"],
                 weight_decay=CFG["weight_decay"], grad_clip=1.0, amp=True, cw=ctx.class_weight,
                 n_train_total=ctx.n_train_total, grad_reducer=ctx.reduce_grads, health_check=ctx.check)
step.run(); step.run()
step.capture(warmup=3)
for _ in range(5):
    step.run()
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    step.run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 50
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step.run()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in evs:
        agg[e.name[:80]][0] += 1
        agg[e.name[:80]][1] += e.device_time
    tot = sum(v[1] for v in agg.values()) / 5
    print(f"# world {world}: {ms:.4f} ms/step (graph), {tot:.1f} us of kernels per step, p2p={ctx.p2p}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if "p2p" in k or "exchange" in k or "nccl" in k.lower():
            print(f"{v[1] / 5:9.1f} us {v[0] // 5:4d}x  {k}")
step.loss_value()
torch.cuda.synchronize(); dist.barrier()
sys.stdout.flush()
os._exit(0)
