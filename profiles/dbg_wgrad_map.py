import sys
sys.path.insert(0, "/root/repo")
import torch
import egnn_b200 as E
from egnn_b200 import ops
from egnn_b200 import synthetic
gr = synthetic.make_elliptic_like(n_nodes=6000, n_edges=7000, n_feats=166, n_timesteps=12, seed=3, hub_degree=200,
                                  t_train_end=8, t_val_end=10, train_window_k=6)
orig = ops._gemm
def dbg(A, a_sm, a_sk, B, b_sk, b_sn, C, M, N, K, *a, **k):
    try:
        return orig(A, a_sm, a_sk, B, b_sk, b_sn, C, M, N, K, *a, **k)
    except RuntimeError as ex:
        print("FAIL", str(ex)[:60], "A", hex(A.data_ptr()), A.shape, A.stride(), A.storage_offset(), "B", hex(B.data_ptr()), B.shape,
              B.stride(), B.storage_offset(), "C", hex(C.data_ptr()), M, N, K, a_sm, a_sk, b_sk, b_sn, k)
        raise
ops._gemm = dbg
conv = E.SAGEConv(166, 64).cuda()
x = torch.randn(6000, 166).cuda().requires_grad_(True)
y = conv(x, gr.edge_index.cuda())
try:
    y.backward(torch.randn_like(y))
    print("conv backward ok")
except RuntimeError as ex:
    print("conv backward FAIL")
