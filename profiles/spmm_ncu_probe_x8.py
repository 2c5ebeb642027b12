"""ncu target: the three mean-aggregation shapes on the 8x replicated graph (working set far beyond the 126 MB L2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egnn_b200 as E
from egnn_b200 import synthetic, ops, _lib
torch.cuda.set_device(0)
gr = synthetic.replicate(synthetic.make_elliptic_like(train_window_k=8), 8)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
N = gr.num_nodes
g = E.build_graph(ei, N)
torch.manual_seed(0)
only = int(sys.argv[1]) if len(sys.argv) > 1 else 0      # one shape per ncu capture
for F, di, do in ((168, torch.float32, torch.bfloat16), (128, torch.bfloat16, torch.bfloat16), (64, torch.bfloat16, torch.bfloat16)):
    if only and F != only:
        continue
    x = torch.randn(N, F, device='cuda').to(di)
    out = torch.empty(N, F, device='cuda', dtype=do)
    for i in range(2):
        ops.spmm(g, 'csr', _lib.SPMM_MEAN, x, do, out=out)
torch.cuda.synchronize()
print("ok")
