import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import egnn_b200 as E
from egnn_b200 import ops, synthetic, _lib
from egnn_b200.train import TrainStep
gr = synthetic.make_elliptic_like(train_window_k=8)
cfg = dict(arch="sage", in_dim=167, hidden_dim=128, layers=3, dropout=0.4, sym=True, lr=1e-3, wd=5e-5)
x = torch.cat([gr.x, (gr.timestep.float() / float(gr.timestep.max())).unsqueeze(1)], dim=1)
ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
torch.manual_seed(0)
model = E.build_model("sage", 167, cfg).cuda()
model.set_dropout_seed(2024)
orig = ops.linear_wgrad
def spy(g, xx, impl=None):
    out = orig(g, xx, impl=impl)
    if g.dtype == torch.float32 and g.size(0) > 1000:
        ref = g.double().t() @ xx.double()
        simt = orig(g, xx, impl=1)
        ea = float((out.double() - ref).abs().max() / ref.abs().max())
        es = float((simt.double() - ref).abs().max() / ref.abs().max())
        amp = float((g.double().abs().t() @ xx.double().abs()).max() / ref.abs().max())
        print(f"wgrad G{tuple(g.shape)} X{tuple(xx.shape)} ld={xx.stride(0)}: tf32x3 err {ea:.2e} simt err {es:.2e} "
              f"cancellation (sum|terms| / |sum|) {amp:.1f}  g absmax {float(g.abs().max()):.2e} nonzero rows {int((g.abs().sum(1) > 0).sum())}")
    return out
ops.linear_wgrad = spy
step = TrainStep(model, x.cuda(), ei.cuda(), gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(), lr=1e-3, weight_decay=5e-5, amp=False)
step.run()
