"""ncu target: four eager rec_k8 train steps (bf16 autocast), no torch profiler (CUPTI and ncu do not share)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import egnn_b200 as E
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
for _ in range(4):
    step.run()
torch.cuda.synchronize()
print("ok", float(step.loss))
