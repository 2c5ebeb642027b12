"""Epoch tail on the device (csrc/metrics.cu) against the oracle: integer counts exact, AP within 1e-12 relative
(float64; the only difference is the summation order of at most `thresholds` terms)."""
import numpy as np
import pytest
import torch

from oracle import metrics_np as M

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _inputs(n, seed, pos_rate=0.1, quant=0, unknown=0.5):
    r = np.random.default_rng(seed)
    y = np.where(r.random(n) < unknown, -1, (r.random(n) < pos_rate).astype(np.int64)).astype(np.int64)
    s = r.random(n).astype(np.float32)
    if quant:
        s = (np.round(s * quant) / quant).astype(np.float32)
    mask = (y >= 0) & (r.random(n) < 0.6)
    return y, s, mask


@pytest.mark.parametrize("n,quant", [(1, 0), (7, 2), (1000, 0), (2049, 8), (46564, 0), (203769, 0), (203769, 64),
                                     (1 << 20, 1024)])
def test_average_precision_matches_oracle(egnn, n, quant):
    from egnn_b200 import metrics
    y, s, mask = _inputs(n, seed=n + quant, quant=quant)
    out = metrics.average_precision(torch.from_numpy(y).cuda(), torch.from_numpy(mask).cuda(),
                                    scores=torch.from_numpy(s).cuda()).cpu().numpy()
    ap, cnt, pos, thr = M.average_precision((y[mask] == 1).astype(int), s[mask])
    assert (int(out[1]), int(out[2]), int(out[3])) == (cnt, pos, thr)
    assert out[0] == pytest.approx(ap, rel=TOL, abs=1e-15)
    roc = M.roc_auc((y[mask] == 1).astype(int), s[mask])
    assert (np.isnan(roc) and np.isnan(out[4])) or out[4] == pytest.approx(roc, rel=TOL)
    # determinism: bitwise identical on a second run
    out2 = metrics.average_precision(torch.from_numpy(y).cuda(), torch.from_numpy(mask).cuda(),
                                     scores=torch.from_numpy(s).cuda()).cpu().numpy()
    assert np.array_equal(out, out2, equal_nan=True)


def test_degenerate_selections(egnn):
    from egnn_b200 import metrics
    y = torch.tensor([1, 0, 0, 1, 0], device="cuda")
    same = torch.full((5,), 0.5, device="cuda")
    r = metrics.average_precision(y, None, scores=same).cpu().tolist()
    assert r[:4] == [pytest.approx(0.4), 5.0, 2.0, 1.0] and r[4] == pytest.approx(0.5)   # one threshold: chance level
    none = torch.zeros(5, dtype=torch.bool, device="cuda")
    r = metrics.average_precision(y, none, scores=same).cpu().tolist()
    assert r[:4] == [0.0, 0.0, 0.0, 0.0] and r[4] != r[4]                                # ROC-AUC undefined: NaN
    neg = torch.zeros(5, dtype=torch.int64, device="cuda")
    assert metrics.average_precision(neg, None, scores=torch.arange(5, device="cuda").float()).cpu().tolist()[:4] == \
        [0.0, 5.0, 0.0, 5.0]
    with pytest.raises(RuntimeError):
        metrics.average_precision(y.cpu(), None, scores=same.cpu())      # no CPU fallback


def test_logits_path_matches_eval_split(egnn):
    """score = softmax(logits)[:, 1] (src/train_gnn.py:254); PR-AUC identical to 3 decimals and far better."""
    from egnn_b200 import metrics
    g = torch.Generator().manual_seed(5)
    n = 30000
    logits = torch.randn(n, 2, generator=g) * 3
    y = (torch.rand(n, generator=g) < 0.1).long()
    mask = torch.rand(n, generator=g) < 0.5
    probs = torch.softmax(logits, dim=1)[:, 1].numpy()
    want = M.average_precision((y.numpy()[mask.numpy()] == 1).astype(int), probs[mask.numpy()])
    sc = torch.empty(n, device="cuda")
    out = metrics.average_precision(y.cuda(), mask.cuda(), logits=logits.cuda(), scores_out=sc).cpu().numpy()
    assert np.abs(sc.cpu().numpy() - probs).max() <= 2e-7
    assert (int(out[1]), int(out[2])) == (want[1], want[2])
    assert abs(out[0] - want[0]) < 1e-5


def test_early_stopper_matches_reference_bookkeeping(egnn):
    from egnn_b200 import metrics
    flat = torch.arange(1003, dtype=torch.float32, device="cuda")
    es = metrics.EarlyStopper(patience=3, flat_param=flat)
    ref = M.EarlyStop()
    snap = None
    for i, v in enumerate([0.2, 0.3, 0.3, 0.25, 0.31, 0.1, 0.1, 0.1]):
        flat += 1.0                                    # "training" changes the parameters every epoch
        es.update(torch.tensor([v, 0, 0, 0], dtype=torch.float64, device="cuda"))
        if ref.update(v):
            snap = flat.clone()
        st = es.state.cpu().tolist()
        assert st[:4] == [ref.best, float(ref.bad), float(ref.best_epoch), float(ref.epoch)]
        assert torch.equal(es.best_param, snap)
    assert es.should_stop() and es.best == 0.31
    es.restore_best()
    assert torch.equal(flat, snap)


def test_fit_loop_matches_host_side_early_stopping(egnn):
    """metrics.fit (device-side epoch tail, polled every 4 epochs) against the reference's loop shape
    (src/train_gnn.py:380-417) run on the host with the oracle metric: same best epoch, same best value to 1e-9,
    and the restored parameters / BatchNorm buffers are the ones of the best epoch."""
    from egnn_b200 import metrics, synthetic
    from egnn_b200.train import TrainStep, eval_probs
    cfg = dict(hidden_dim=32, layers=3, dropout=0.0, time_embed_dim=2, time_embed_type="sin", max_timestep=49)
    gr = synthetic.make_elliptic_like(n_nodes=5000, n_edges=6000, n_timesteps=10, seed=4, hub_degree=80,
                                      t_train_end=6, t_val_end=8)
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1).cuda()
    x, t, y = gr.x.cuda(), gr.timestep.cuda(), gr.y.cuda()
    tm, vm = gr.train_mask.cuda(), gr.val_mask.cuda()
    kw = dict(lr=3e-2, weight_decay=0.0, grad_clip=1.0, amp=False)
    patience, max_epochs = 3, 24

    torch.manual_seed(0)
    ref_model = egnn.build_model("sage_resbn", 166, cfg).cuda()
    step = TrainStep(ref_model, x, ei, t, y, tm, **kw)
    step.run()                                            # metrics.fit runs one eager step before its loop, too
    es, best_state, stop_epoch = M.EarlyStop(), None, None
    yv = (gr.y.numpy()[gr.val_mask.numpy()] == 1).astype(int)
    for epoch in range(1, max_epochs + 1):
        step.run()
        probs, _ = eval_probs(ref_model, x, ei, t)
        ap = M.average_precision(yv, probs.cpu().numpy()[gr.val_mask.numpy()])[0]
        if es.update(ap):
            best_state = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
        if es.bad >= patience:
            stop_epoch = epoch
            break
    if stop_epoch is None:
        stop_epoch = max_epochs          # never triggered: both loops run to the end

    torch.manual_seed(0)
    model = egnn.build_model("sage_resbn", 166, cfg).cuda()
    res = metrics.fit(model, x, ei, t, y, tm, vm, max_epochs=max_epochs, patience=patience, poll_every=4,
                      capture=False, **kw)
    assert res["best_epoch"] == es.best_epoch
    assert res["best_val"] == pytest.approx(es.best, rel=1e-6)   # device softmax vs torch.softmax: last-ulp scores
    assert stop_epoch <= res["epochs"] < stop_epoch + 4 and res["epochs"] <= max_epochs   # bounded overshoot
    got = model.state_dict()
    for k, v in best_state.items():
        assert torch.equal(got[k], v), k
