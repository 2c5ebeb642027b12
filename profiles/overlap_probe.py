"""rec_k8 graphed step with / without the weight-gradient GEMMs on a side stream (fused.OVERLAP_WGRAD) at N=1."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import egnn_b200 as E
from egnn_b200 import fused
from egnn_b200.train import TrainStep
torch.cuda.set_device(0)
gr = bench.host_graph(1)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
x, t, y, m = gr.x.cuda(), gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda()
CFG = dict(bench.CFG)
for mode in ("auto", 1, "auto", 1):
    fused.OVERLAP_WGRAD = mode
    torch.manual_seed(42)
    model = E.build_model(CFG["arch"], gr.x.size(1), CFG).cuda()
    model.set_dropout_seed(42, "cuda")
    step = TrainStep(model, x, ei, t, y, m, lr=CFG["lr"], weight_decay=CFG["weight_decay"], grad_clip=1.0, amp=True)
    step.run()
    step.capture(warmup=2)
    for _ in range(5):
        step.run()
    ms = bench._timed(step.run, 100)
    print(f"OVERLAP_WGRAD={mode}: {ms * 1e3:.1f} us/step, loss {float(step.loss):.6f}", flush=True)
