#!/usr/bin/env python
"""N-GPU check that the timestep-sharded step equals the single-GPU step on the SAME graph (rec_k8, fp32 and
bf16, eager and CUDA-graph replays): identical initial weights, 3 steps; compares the loss trajectory, the all-reduced
step-1 gradients and the final parameters.  Tensors whose gradient is ANALYTICALLY ZERO (the conv bias in front of a
BatchNorm: BN subtracts the column mean, so d loss / d bias == 0 and both sides hold rounding noise that Adam turns
into +-lr steps) are excluded by name from the gradient / parameter comparison, not waved through.
launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/shard_check.py
(tests/test_gpu_multi.py runs exactly this when two GPUs are visible)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import egnn_b200 as E
from egnn_b200 import synthetic
from egnn_b200.shard import ShardedContext, make_shard
from egnn_b200.train import TrainStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
CFG = dict(bench.CFG)
gr = synthetic.make_elliptic_like(n_nodes=60000, n_edges=69000, n_timesteps=20, seed=11, hub_degree=300,
                                  t_train_end=14, t_val_end=17, train_window_k=8)


def make_step(graph, ctx, amp):
    torch.manual_seed(7)
    model = E.build_model(CFG["arch"], graph.x.size(1), dict(CFG, max_timestep=49)).to(dev)
    model.set_dropout_seed(123, dev)
    if ctx is not None:
        ctx.attach(model)
    ei = torch.cat([graph.edge_index, graph.edge_index.flip(0)], 1).to(dev)
    kw = dict(cw=ctx.class_weight, n_train_total=ctx.n_train_total, grad_reducer=ctx.reduce_grads) if ctx else {}
    return model, TrainStep(model, graph.x.to(dev), ei, graph.timestep.to(dev), graph.y.to(dev),
                            graph.train_mask.to(dev), lr=CFG["lr"], weight_decay=CFG["weight_decay"], grad_clip=1.0,
                            amp=amp, **kw)


def zero_grad_names(model):
    """conv biases feeding BatchNorm (hidden layers of SAGE-ResBN with use_bn)"""
    n_hidden = len(model.convs) - 1
    return {f"convs.{i}.lin_l.bias" for i in range(n_hidden)} if model.use_bn else set()


ok = True
lines = []
for amp, graphed in ((False, False), (True, False), (True, True)):
    sh = make_shard(gr, rank, world)
    ctx = ShardedContext(sh, dev)
    m_sh, st_sh = make_step(sh.graph, ctx, amp)
    st_sh.health_check = ctx.check
    if graphed:
        st_sh.capture(warmup=2, preserve_state=True)
    losses, g1 = [], None
    for it in range(3):
        st_sh.run()
        if it == 0:
            g1 = st_sh.opt.flat_grad.detach().clone()          # all-reduced gradient of step 1
        l = st_sh.loss.detach().clone()           # local loss sum / global count
        dist.all_reduce(l)
        losses.append(float(l))
    st_sh.loss_value()                            # host sync + ctx.check(): raises if a peer-memory exchange timed out
    if rank == 0:
        m_1, st_1 = make_step(gr, None, amp)
        ref, g1_ref = [], None
        for it in range(3):
            st_1.run()
            if it == 0:
                g1_ref = st_1.opt.flat_grad.detach().clone()
            ref.append(float(st_1.loss))
        skip = zero_grad_names(m_1)
        off, gd, pd, gd_name = 0, 0.0, 0.0, ""
        gmax = float(g1_ref.abs().max())
        for (n, a), (_, b) in zip(m_sh.named_parameters(), m_1.named_parameters()):
            k = b.numel()
            ga, gb = g1[off:off + k], g1_ref[off:off + k]
            off += k
            if n in skip:
                # rounding noise of a cancelling sum (bf16: the sum of ~2e5 rounded dz values)
                assert float(gb.abs().max()) < (5e-2 if amp else 1e-4) * gmax, (n, "expected an analytically zero gradient",
                                                                               float(gb.abs().max()), gmax)
                continue
            d_g = float((ga - gb).abs().max() / gb.abs().max().clamp_min(1e-30))
            d_p = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
            if d_g > gd:
                gd, gd_name = d_g, n
            pd = max(pd, d_p)
        tol_l, tol_g, tol_p = (2e-2, 5e-2, 5e-2) if amp else (1e-5, 2e-4, 5e-4)
        good = all(abs(a - b) <= tol_l * abs(b) for a, b in zip(losses, ref)) and gd <= tol_g and pd <= tol_p
        ok &= good
        line = (f"{'bf16' if amp else 'fp32'}{' cuda-graph' if graphed else ''} world={world} p2p={ctx.p2p}: sharded losses "
                f"{losses} | single-GPU {ref} | worst step-1 gradient rel diff {gd:.2e} ({gd_name}) | worst param rel diff after 3 "
                f"steps {pd:.2e} (excluded analytically-zero: {sorted(skip)}) -> {'OK' if good else 'MISMATCH'}")
        lines.append(line)
        print(line, flush=True)
    dist.barrier()
if rank == 0:
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02")
    if os.path.isdir(out):
        open(os.path.join(out, f"shard_check_{world}gpu.txt"), "w").write("\n".join(lines) + "\n")
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if ok else 1)
