"""Generate tests/golden/*.pt by running the REFERENCE's own code (unmodified, imported from
/root/reference) wherever it can run here.

torch_geometric is not installable in this image, so `src/models/gnn.py` cannot be imported
as is.  This script registers stub `torch_geometric.*` modules whose GCNConv/SAGEConv/GATConv
are the restated convs of oracle/pyg_restated.py, then imports the reference's own
  * src.models.gnn    (GCNNet / SAGENet / GATNet / SAGEResBNNet composition, _sinusoid ...)
  * src.train_gnn     (build_model, class_weight, _make_loss_fn, train_epoch, eval_split)
  * src.data.dataset_elliptic.make_temporal_masks
and records their outputs on a small seeded Elliptic-shaped graph.  What this pins: the
net-level composition, the loss, clip + Adam step and the temporal masks are the reference's
own code; the conv arithmetic itself remains the restatement (PARITY UNPINNED there).

Run in the build container only:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import pyg_restated as O  # noqa: E402


class Data:  # stand-in for torch_geometric.data.Data (attribute bag with .to())
    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def num_nodes(self):
        return self.x.size(0)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


def _stub():
    tg = types.ModuleType("torch_geometric")
    nn_ = types.ModuleType("torch_geometric.nn")
    nn_.GCNConv, nn_.SAGEConv, nn_.GATConv = O.GCNConv, O.SAGEConv, O.GATConv
    ld = types.ModuleType("torch_geometric.loader")
    ld.NeighborLoader = object
    dt = types.ModuleType("torch_geometric.data")
    dt.Data = Data
    tg.nn, tg.loader, tg.data = nn_, ld, dt
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": nn_, "torch_geometric.loader": ld,
                        "torch_geometric.data": dt})


CONFIGS = {
    "gcn": dict(arch="gcn", hidden_dim=16, layers=3, dropout=0.0, lr=3e-3, weight_decay=1e-4, grad_clip=1.0,
                symmetrize_edges=False, use_time_scalar=True, train_window_k=4),
    "sage": dict(arch="sage", hidden_dim=16, layers=2, dropout=0.0, lr=3e-3, weight_decay=1e-4, grad_clip=1.0,
                 symmetrize_edges=True, use_time_scalar=True, train_window_k=4),
    "rec_k8": dict(arch="sage_resbn", hidden_dim=16, layers=3, dropout=0.0, lr=5e-4, weight_decay=5e-5,
                   grad_clip=1.0, symmetrize_edges=True, use_time_scalar=False, time_embed_dim=2,
                   time_embed_type="sin", max_timestep=49, train_window_k=3),
    "gat": dict(arch="gat", hidden_dim=16, layers=2, heads=4, dropout=0.0, lr=3e-3, weight_decay=1e-4,
                grad_clip=1.0, symmetrize_edges=False, use_time_scalar=True, train_window_k=4),
}


def main():
    _stub()
    from src.data.dataset_elliptic import make_temporal_masks
    from src import train_gnn as T
    from egnn_b200 import synthetic

    gr = synthetic.make_elliptic_like(n_nodes=600, n_edges=700, n_feats=22, n_timesteps=8, seed=11,
                                      hub_degree=80)
    out = {}
    for name, cfg in CONFIGS.items():
        torch.manual_seed(42)
        data = Data(x=gr.x.clone(), edge_index=gr.edge_index.clone(), y=gr.y.clone())
        data.timestep = gr.timestep.clone()
        data = make_temporal_masks(data, t_train_end=5, t_val_end=6, train_window_k=cfg["train_window_k"])
        if cfg.get("use_time_scalar", False) and cfg.get("time_embed_dim", 0) == 0:  # train_gnn.py:315-317
            tnorm = (data.timestep.float() / float(data.timestep.max())).unsqueeze(1)
            data.x = torch.cat([data.x, tnorm], dim=1)
        ei = data.edge_index
        if cfg.get("symmetrize_edges", False):                                      # train_gnn.py:321-324
            ei = torch.cat([ei, ei.flip(0)], dim=1)
        model = T.build_model(cfg["arch"], data.x.size(1), cfg)
        state0 = {k: v.clone() for k, v in model.state_dict().items()}
        opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
        cw = T.class_weight(data.y[data.train_mask])
        t_train = data.timestep[data.train_mask]
        loss_fn = T._make_loss_fn(cfg, cw, model, int(t_train.min()), int(t_train.max()))
        dev = torch.device("cpu")
        scaler = T._make_grad_scaler(dev, False)
        _, _, logits_eval0 = T.eval_split(model, data, ei, data.val_mask)
        losses = [T.train_epoch(model, data, ei, opt, loss_fn, scaler, False, cfg, dev) for _ in range(2)]
        _, p_val, logits_eval2 = T.eval_split(model, data, ei, data.val_mask)
        out[name] = dict(cfg=cfg, x=data.x, edge_index_used=ei, y=data.y, timestep=data.timestep,
                         train_mask=data.train_mask, val_mask=data.val_mask, test_mask=data.test_mask,
                         class_weight=cw, state0=state0, logits_eval0=logits_eval0.clone(), losses=losses,
                         state2={k: v.clone() for k, v in model.state_dict().items()},
                         logits_eval2=logits_eval2.clone())
        print(name, "losses", losses, "params", sum(p.numel() for p in model.parameters()))
    torch.save(out, os.path.join(HERE, "reference_nets.pt"))
    print("wrote", os.path.join(HERE, "reference_nets.pt"), os.path.getsize(os.path.join(HERE, "reference_nets.pt")))


if __name__ == "__main__":
    main()
