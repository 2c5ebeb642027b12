"""Generate tests/golden/minibatch_golden.pt by running the REFERENCE's own mini-batch loops, unmodified, imported from
/root/reference (`src.train_gnn.train_epoch_minibatch`, `:212-245`; `eval_val_minibatch`, `:261-280`; with its own
`build_model`, `_make_loss_fn`, `_make_grad_scaler`), on batches drawn by the sequential sampler oracle
(`oracle/neighbor_sample_np.py` -- PyG's NeighborLoader itself is not installable).  torch_geometric is
stubbed as in tests/golden/make_golden.py (restated convs).  What this pins: the loop around the model -- seeds-first
slicing, the time-weighting switch, clip + Adam per batch, the seed-weighted epoch loss, the validation concatenation.

Run in the build container only:  python tests/golden/make_minibatch_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden as MG  # noqa: E402

MG._stub()
from src import train_gnn as T  # noqa: E402
from src.data.dataset_elliptic import make_temporal_masks  # noqa: E402

from egnn_b200 import synthetic  # noqa: E402
from oracle.neighbor_sample_np import csc_by_destination, neighbor_sample  # noqa: E402

CONFIGS = {
    "sage": dict(arch="sage", hidden_dim=16, layers=2, dropout=0.0, lr=3e-3, weight_decay=1e-4, grad_clip=1.0),
    "rec_k8_timeweighted": dict(arch="sage_resbn", hidden_dim=16, layers=3, dropout=0.0, lr=5e-4, weight_decay=5e-5,
                                grad_clip=0.5, time_embed_dim=2, time_embed_type="sin", max_timestep=49,
                                time_loss_weighting="linear"),
}


class Batch(MG.Data):
    pass


def make_batches(data, ei, input_nodes, fanouts, batch_size, seed):
    ip, src, eid = csc_by_destination(ei.numpy(), data.x.size(0))
    out = []
    for b, lo in enumerate(range(0, input_nodes.numel(), batch_size)):
        seeds = input_nodes[lo:lo + batch_size]
        n_id, le, _, _, _ = neighbor_sample(ip, src, eid, seeds.tolist(), fanouts, seed=seed, batch_idx=b)
        n_id = torch.as_tensor(np.asarray(n_id), dtype=torch.long)
        out.append(Batch(x=data.x[n_id], y=data.y[n_id], timestep=data.timestep[n_id], n_id=n_id,
                         edge_index=torch.as_tensor(np.asarray(le), dtype=torch.long).view(2, -1),
                         batch_size=int(seeds.numel())))
    return out


def main():
    gr = synthetic.make_elliptic_like(n_nodes=500, n_edges=650, n_feats=12, n_timesteps=8, seed=19, hub_degree=60)
    out = {}
    dev = torch.device("cpu")
    for name, cfg in CONFIGS.items():
        torch.manual_seed(7)
        data = MG.Data(x=gr.x.clone(), edge_index=gr.edge_index.clone(), y=gr.y.clone())
        data.timestep = gr.timestep.clone()
        data = make_temporal_masks(data, t_train_end=5, t_val_end=6, train_window_k=None)
        ei = torch.cat([data.edge_index, data.edge_index.flip(0)], dim=1)
        train_idx = torch.nonzero(data.train_mask).view(-1)
        val_idx = torch.nonzero(data.val_mask).view(-1)
        train_batches = make_batches(data, ei, train_idx, [3, 2], 20, seed=3)
        val_batches = make_batches(data, ei, val_idx, [4, 4], 5, seed=4)
        model = T.build_model(cfg["arch"], data.x.size(1), cfg)
        state0 = {k: v.clone() for k, v in model.state_dict().items()}
        opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
        cw = T.class_weight(data.y[data.train_mask])
        t_train = data.timestep[data.train_mask]
        t_min, t_max = int(t_train.min()), int(t_train.max())
        loss_fn = T._make_loss_fn(cfg, cw, model, t_min, t_max)
        scaler = T._make_grad_scaler(dev, False)
        losses = [T.train_epoch_minibatch(model, train_batches, opt, loss_fn, scaler, False, cfg, dev)
                  for _ in range(2)]
        y_val, p_val = T.eval_val_minibatch(model, val_batches, dev)
        keep = lambda bs: [dict(n_id=b.n_id.to(torch.int32), edge_index=b.edge_index.to(torch.int32),
                                batch_size=b.batch_size) for b in bs]
        out[name] = dict(cfg=cfg, x=data.x, y=data.y, timestep=data.timestep, class_weight=cw, t_min=t_min, t_max=t_max,
                         train_batches=keep(train_batches), val_batches=keep(val_batches), state0=state0,
                         losses=losses, state2={k: v.clone() for k, v in model.state_dict().items()},
                         y_val=torch.as_tensor(y_val), p_val=torch.as_tensor(p_val))
        print(name, "epoch losses", losses, "batches", len(train_batches), len(val_batches), "val rows", len(y_val))
    path = os.path.join(HERE, "minibatch_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
