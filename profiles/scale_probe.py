#!/usr/bin/env python
"""Strong scaling of BASELINE config 5 (3-layer SAGE, configs/sage_l3_k18.yaml) on the 64x replicated Elliptic-shaped
graph: the 64 x 49 (replica, timestep) blocks are split over the ranks as contiguous replica ranges, zero halo; every
rank generates only its own replicas (same seeds as synthetic.replicate), so the GLOBAL graph is identical for every N.
launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/scale_probe.py [--replicas 64]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import egnn_b200 as E
from egnn_b200 import synthetic
from egnn_b200.shard import Shard, ShardedContext
from egnn_b200.train import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    out_fd = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
cfg = dict(arch="sage", hidden_dim=128, layers=3, dropout=0.4, lr=1e-3, weight_decay=5e-4)
base = synthetic.make_elliptic_like(train_window_k=18)
n1, k = base.num_nodes, args.replicas
assert k % world == 0, "replicas must divide evenly over the ranks"
r0, r1 = rank * (k // world), (rank + 1) * (k // world)
xs = []
for r in range(r0, r1):   # same per-replica feature streams as synthetic.replicate(base, k)
    xs.append(base.x if r == 0 else synthetic._features(n1, base.x.size(1), torch.Generator().manual_seed(42 + r)))
x = torch.cat(xs)
kk = r1 - r0
tstep = base.timestep.repeat(kk)
x = torch.cat([x, (tstep.float() / tstep.max().float()).unsqueeze(1)], dim=1)      # use_time_scalar
ei = torch.cat([base.edge_index + j * n1 for j in range(kk)], dim=1)
local_g = synthetic.EllipticGraph(x=x, edge_index=ei, y=base.y.repeat(kk), timestep=tstep,
                                  train_mask=base.train_mask.repeat(kk), val_mask=base.val_mask.repeat(kk),
                                  test_mask=base.test_mask.repeat(kk))
sh = Shard(rank=rank, world=world, row0=r0 * n1, n_local=kk * n1, n_global=k * n1, graph=local_g)
ctx = ShardedContext(sh, dev)
torch.manual_seed(42)
model = ctx.attach(E.build_model(cfg["arch"], x.size(1), cfg).to(dev))
model.set_dropout_seed(42, dev)
ei_sym = torch.cat([ei, ei.flip(0)], 1).to(dev)
step = TrainStep(model, x.to(dev), ei_sym, tstep.to(dev), local_g.y.to(dev), local_g.train_mask.to(dev), lr=cfg["lr"],
                 weight_decay=cfg["weight_decay"], grad_clip=1.0, amp=True, cw=ctx.class_weight,
                 n_train_total=ctx.n_train_total, grad_reducer=ctx.reduce_grads if world > 1 else None)
e_total = 2 * base.edge_index.size(1) * k
del x, xs, local_g
step.run()
torch.cuda.synchronize()
step.capture(warmup=2)
for _ in range(2):
    step.run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps):
    step.run()
b.record()
torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
loss = step.loss.detach().clone().double()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(loss)
if rank == 0:
    if world > 1:
        sys.stdout.flush()
        os.dup2(out_fd, 1)
    ms = float(t)
    print(f"sage_l3 x{k} on {world} GPU(s): {ms:.3f} ms/step (max over ranks), {e_total / ms / 1e6:.3f} GEdges/s, "
          f"N={k * n1} E'={e_total}, loss {float(loss):.4f}, peak mem/GPU {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, "
          f"collectives: {'peer-memory kernel' if ctx.p2p else ('nccl' if world > 1 else 'none')}", flush=True)
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
sys.stdout.flush()
os._exit(0)
