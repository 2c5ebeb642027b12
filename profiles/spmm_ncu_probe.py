"""Short program for ncu captures of the aggregation kernel alone: the two rec_k8 shapes (layer 0: F=168
fp32 -> bf16; hidden layers: F=64 bf16), mean aggregation over the symmetrised Elliptic-shaped graph."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egnn_b200 as E
from egnn_b200 import synthetic, ops, _lib
torch.cuda.set_device(0)
gr = synthetic.make_elliptic_like(train_window_k=8)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
N = gr.num_nodes
g = E.build_graph(ei, N)
torch.manual_seed(0)
for F, di, do in ((168, torch.float32, torch.bfloat16), (64, torch.bfloat16, torch.bfloat16)):
    xs = [torch.randn(N, F, device='cuda').to(di) for _ in range(3)]
    out = torch.empty(N, F, device='cuda', dtype=do)
    for i in range(3):
        ops.spmm(g, 'csr', _lib.SPMM_MEAN, xs[i], do, out=out)
torch.cuda.synchronize()
print("ok")
