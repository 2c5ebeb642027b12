"""CPU tests pinning the oracle's mini-batch loops (`oracle/pyg_restated.py:train_epoch_minibatch`,
`eval_val_minibatch`) to the REFERENCE's own functions: `tests/golden/minibatch_golden.pt` holds what
`src.train_gnn.train_epoch_minibatch` / `eval_val_minibatch` (unmodified, imported from /root/reference by
`tests/golden/make_minibatch_golden.py`) produce over two epochs on batches drawn by the sequential sampler oracle.
The GPU test (`tests/test_gpu_loader.py`) holds `egnn_b200.train.train_epoch_minibatch` to this same loop."""
import os
import types

import pytest
import torch

from oracle import pyg_restated as O

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "minibatch_golden.pt"))


def _batches(rec, key):
    out = []
    for b in rec[key]:
        n_id = b["n_id"].long()
        out.append(types.SimpleNamespace(x=rec["x"][n_id], y=rec["y"][n_id], timestep=rec["timestep"][n_id],
                                         edge_index=b["edge_index"].long(), batch_size=b["batch_size"]))
    return out


@pytest.mark.parametrize("name", list(GOLD))
def test_minibatch_loops_reproduce_the_reference(name):
    rec = GOLD[name]
    cfg = rec["cfg"]
    model = O.build_model(cfg["arch"], rec["x"].size(1), cfg)
    model.load_state_dict(rec["state0"])
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    loss_fn = O.make_loss_fn(cfg, rec["class_weight"], model, rec["t_min"], rec["t_max"])
    train, val = _batches(rec, "train_batches"), _batches(rec, "val_batches")
    assert len(train) >= 3 and len(val) >= 2 and train[-1].batch_size != train[0].batch_size   # ragged last batch
    losses = [O.train_epoch_minibatch(model, train, opt, loss_fn, cfg) for _ in range(2)]
    for got, want in zip(losses, rec["losses"]):
        assert abs(got - want) <= 1e-6 * abs(want), (got, want)
    for k, v in model.state_dict().items():
        torch.testing.assert_close(v, rec["state2"][k], rtol=1e-6, atol=1e-7, msg=k)
    y, p = O.eval_val_minibatch(model, val)
    assert torch.equal(y, rec["y_val"].long())
    torch.testing.assert_close(p, rec["p_val"].float(), rtol=1e-6, atol=1e-7)


def test_minibatch_loops_edge_cases():
    """No batches: epoch loss 0.0 and empty validation arrays, as the reference returns (`:244-245`, `:276-277`)."""
    model = O.build_model("sage", 4, dict(hidden_dim=4, layers=2, dropout=0.0))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    assert O.train_epoch_minibatch(model, [], opt, lambda *a: None, {}) == 0.0
    y, p = O.eval_val_minibatch(model, [])
    assert y.numel() == 0 and p.numel() == 0
