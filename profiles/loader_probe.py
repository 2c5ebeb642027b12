"""Mini-batch path timing: one NeighborLoader batch (sampling + relabelling + row slices) and one mini-batch training epoch
of rec_k8 with the reference's defaults (fanout [10, 10], batch_size 8192, src/train_gnn.py:333-334) on the full graph.
usage: python profiles/loader_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egnn_b200 as E
from egnn_b200 import synthetic
from egnn_b200.train import train_epoch_minibatch

torch.cuda.set_device(0)
gr = synthetic.make_elliptic_like(train_window_k=8)


class Data:
    pass


data = Data()
data.x, data.y, data.timestep = gr.x.cuda(), gr.y.cuda(), gr.timestep.cuda()
data.train_mask, data.val_mask, data.test_mask = gr.train_mask.cuda(), gr.val_mask.cuda(), gr.test_mask.cuda()
data.edge_index = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
idx = torch.nonzero(data.train_mask).view(-1)
print(f"# N={gr.num_nodes} E={data.edge_index.size(1)} train nodes={idx.numel()}")
for fan, bs in (([10, 10], 8192), ([25, 10], 8192), ([-1, -1], 8192), ([10, 10], 1024)):
    loader = E.NeighborLoader(data, num_neighbors=fan, batch_size=bs, input_nodes=idx, shuffle=True, seed=1)
    for b in loader:      # warm-up epoch
        pass
    torch.cuda.synchronize()
    n0 = E._lib.launch_count()
    t0 = time.perf_counter()
    nb = nn = ne = 0
    for b in loader:
        nb, nn, ne = nb + 1, nn + b.num_nodes, ne + b.edge_index.size(1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"fanout {fan} batch_size {bs}: {nb} batches, {dt / nb * 1e3:.3f} ms per batch (wall, one sync per batch), "
          f"{nn // nb} nodes / {ne // nb} edges per batch, {(E._lib.launch_count() - n0) // nb} kernel launches per batch")
cfg = dict(hidden_dim=64, layers=3, dropout=0.2, time_embed_dim=2, time_embed_type="sin", max_timestep=49)
torch.manual_seed(0)
model = E.build_model("sage_resbn", 166, cfg).cuda()
from egnn_b200.train import class_weight
cw = class_weight(data.y[data.train_mask])
loss_fn = E.make_loss_fn({}, cw, model, 1, 49)
opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=5e-5)
loader = E.NeighborLoader(data, num_neighbors=[10, 10], batch_size=8192, input_nodes=idx, shuffle=True, seed=1)
for amp in (False, True):
    train_epoch_minibatch(model, loader, opt, loss_fn, {"grad_clip": 1.0}, use_amp=amp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        loss = train_epoch_minibatch(model, loader, opt, loss_fn, {"grad_clip": 1.0}, use_amp=amp)
    torch.cuda.synchronize()
    print(f"rec_k8 mini-batch epoch ({len(loader)} batches, fanout [10, 10], batch_size 8192, amp={amp}): "
          f"{(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per epoch (eager, wall), loss {loss:.4f}")
