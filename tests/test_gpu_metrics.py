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


def test_early_stopper_freezes_at_patience(egnn):
    """a value that improves AFTER `bad >= patience` (epochs the host runs before it polls) must not move
    the best value or overwrite the snapshot -- the reference has left its loop by then (src/train_gnn.py:411)."""
    from egnn_b200 import metrics
    flat = torch.zeros(64, dtype=torch.float32, device="cuda")
    es = metrics.EarlyStopper(patience=2, flat_param=flat)
    seq = [0.5, 0.4, 0.3, 0.9, 0.95, 0.1]           # reference: best 0.5 at epoch 1, breaks after epoch 3
    for v in seq:
        flat += 1.0
        es.update(torch.tensor([v, 0, 0, 0], dtype=torch.float64, device="cuda"))
    st = es.state.cpu().tolist()
    assert st[:4] == [0.5, 2.0, 1.0, 3.0] and st[4] == 0.0
    assert torch.equal(es.best_param, torch.full((64,), 1.0, device="cuda"))
    es.restore_best()
    assert torch.equal(flat, torch.full((64,), 1.0, device="cuda"))
    # patience <= 0: never frozen (plain bookkeeping)
    es = metrics.EarlyStopper(patience=0, flat_param=flat)
    for v in seq:
        es.update(torch.tensor([v, 0, 0, 0], dtype=torch.float64, device="cuda"))
    assert es.state.cpu().tolist()[:4] == [0.95, 1.0, 5.0, 6.0]


@pytest.mark.parametrize("capture", [False, True])
def test_fit_loop_matches_host_side_early_stopping(egnn, capture):
    """metrics.fit (device-side epoch tail, polled every 4 epochs; eager and CUDA-graph captured) against the
    reference's loop shape (src/train_gnn.py:380-417) run on the host with the oracle metric, both starting from the
    SAME initial weights with epoch 1 = the first optimizer step: same best epoch, same stopping epoch, same best
    value, and the restored parameters / BatchNorm buffers are bitwise the ones of the best epoch."""
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
                      capture=capture, **kw)
    assert res["best_epoch"] == es.best_epoch
    assert res["stop_epoch"] == stop_epoch
    assert res["best_val"] == pytest.approx(es.best, rel=1e-6)   # device softmax vs torch.softmax: last-ulp scores
    assert stop_epoch <= res["epochs"] < stop_epoch + 4 and res["epochs"] <= max_epochs   # bounded overshoot
    got = model.state_dict()
    for k, v in best_state.items():
        assert torch.equal(got[k], v), k


@pytest.mark.parametrize("n,quant", [(1, 0), (9, 2), (1000, 0), (2049, 8), (46564, 0), (203769, 64), (203769, 0)])
def test_ranking_metrics_match_oracle(egnn, n, quant):
    """every number of the reference's run tail (src/train_gnn.py:449-470; src/utils/metrics.py:18-66)
    from one device sort, against the oracle that is pinned to the reference's own functions."""
    from egnn_b200 import metrics
    y, s, mask = _inputs(n, seed=3 * n + quant, quant=quant)
    yb, sb = (y[mask] == 1).astype(int), s[mask]
    k = 100
    thr_in = torch.tensor([0.37], dtype=torch.float64, device="cuda")
    yc, mc, sc = torch.from_numpy(y).cuda(), torch.from_numpy(mask).cuda(), torch.from_numpy(s).cuda()
    out = metrics.ranking_metrics(yc, mc, scores=sc, top_k=k, target_precision=0.2, threshold=thr_in).cpu().numpy()
    out_self = metrics.ranking_metrics(yc, mc, scores=sc, top_k=k, target_precision=0.2).cpu().numpy()
    if yb.size == 0:
        assert out[1] == 0 and out[8] == 0.0
        return
    ap = M.average_precision(yb, sb)
    assert out[0] == pytest.approx(ap[0], rel=TOL, abs=1e-15) and (int(out[1]), int(out[2])) == ap[1:3]
    thr, f1 = M.pick_threshold_max_f1(yb, sb)
    if yb.sum() > 0:
        assert out[9] == thr and out[8] == pytest.approx(f1, rel=1e-12)
        assert out[11] == pytest.approx(M.recall_at_precision(yb, sb, 0.2), rel=1e-12)
        assert out[12] == M.pick_threshold_for_precision(yb, sb, 0.2)
        assert out_self[13] == pytest.approx(M.f1_at_threshold(yb, sb, thr), rel=1e-12)
    assert out[13] == pytest.approx(M.f1_at_threshold(yb, sb, 0.37), rel=1e-12, abs=1e-15)
    assert int(out[15]) == min(k, yb.size)
    if not quant:   # ties at the k-th score are unspecified in the reference (unstable argsort)
        assert out[10] == pytest.approx(M.precision_at_k(yb, sb, k), rel=1e-12, abs=1e-15)
    assert out[14] == pytest.approx(M.expected_calibration_error(yb, sb), rel=1e-6, abs=1e-9)   # fp32 bin means there


def test_temperature_fit_matches_reference_scaler(egnn):
    """egnn_temperature_fit against golden T values from the reference's own TemperatureScaler.fit
    (tests/golden/make_metrics_golden.py) and the oracle minimiser."""
    import json
    import os
    from egnn_b200 import metrics
    from test_oracle_metrics import _temp_case
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_golden.json")))
    for c in g["temperature"]:
        logits, yy = _temp_case(c["seed"], c["n"], c["scale"])
        out = metrics.fit_temperature(logits.cuda(), yy.cuda()).cpu().tolist()
        T_or = M.fit_temperature(logits.numpy(), yy.numpy())
        assert out[0] == pytest.approx(T_or, rel=1e-6), c
        assert out[2] == pytest.approx(M.temperature_nll(logits.numpy(), yy.numpy(), T_or), rel=1e-9)
        assert out[1] == pytest.approx(M.temperature_nll(logits.numpy(), yy.numpy(), 1.0), rel=1e-9)
        assert int(out[4]) == c["n"] and out[3] < 50
        if not c["reference_diverged"]:
            assert out[0] == pytest.approx(c["T"], rel=2e-3), c      # the reference's fp32 LBFGS stops ~1e-3 short
    # masked rows and unlabelled rows (-1) are ignored
    logits, yy = _temp_case(0, 4000, 3.0)
    y2 = yy.clone()
    y2[::3] = -1
    m = torch.rand(4000, generator=torch.Generator().manual_seed(1)) < 0.7
    sel = (m & (y2 >= 0)).numpy()
    out = metrics.fit_temperature(logits.cuda(), y2.cuda(), m.cuda()).cpu().tolist()
    assert int(out[4]) == int(sel.sum())
    assert out[0] == pytest.approx(M.fit_temperature(logits.numpy()[sel], y2.numpy()[sel]), rel=1e-6)


def test_final_metrics_matches_the_reference_run_tail(egnn):
    """metrics.final_metrics (device) against the reference's run tail restated with the oracle functions
    (src/train_gnn.py:420-480): temperature on val, threshold on val, metrics on test."""
    from egnn_b200 import metrics, synthetic
    from egnn_b200.train import TrainStep, eval_probs
    cfg = dict(hidden_dim=32, layers=3, dropout=0.0, time_embed_dim=2, time_embed_type="sin", max_timestep=49)
    gr = synthetic.make_elliptic_like(n_nodes=20000, n_edges=23000, n_timesteps=12, seed=6, hub_degree=80,
                                      t_train_end=7, t_val_end=9, label_signal=1.0)
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1).cuda()
    x, t, y = gr.x.cuda(), gr.timestep.cuda(), gr.y.cuda()
    torch.manual_seed(0)
    model = egnn.build_model("sage_resbn", 166, cfg).cuda()
    step = TrainStep(model, x, ei, t, y, gr.train_mask.cuda(), lr=1e-2, weight_decay=0.0, amp=False)
    for _ in range(15):
        step.run()
    vm, tm = gr.val_mask.numpy(), gr.test_mask.numpy()
    for calib in (False, True):
        got = metrics.final_metrics(model, x, ei, t, y, gr.val_mask.cuda(), gr.test_mask.cuda(),
                                    calibrate_temperature=calib, top_k=50)
        _, logits = eval_probs(model, x, ei, t)
        logits = logits.float().cpu()
        yn = gr.y.numpy()
        if calib:
            T = M.fit_temperature(logits.numpy()[vm], yn[vm])
            assert got["temperature"] == pytest.approx(T, rel=1e-6)
            logits = logits / torch.tensor(got["temperature"], dtype=torch.float32)
        probs = torch.softmax(logits, dim=1)[:, 1].numpy()
        yv, pv, yt, pt = (yn[vm] == 1).astype(int), probs[vm], (yn[tm] == 1).astype(int), probs[tm]
        thr, _ = M.pick_threshold_max_f1(yv, pv)
        # device softmax vs torch.softmax differ in the last ulp of a score: compare at 1e-6
        assert got["threshold"] == pytest.approx(thr, rel=1e-6)
        assert got["pr_auc_illicit"] == pytest.approx(M.average_precision(yt, pt)[0], rel=1e-6)
        assert got["roc_auc"] == pytest.approx(M.roc_auc(yt, pt), rel=1e-6)
        assert got["f1_illicit_at_thr"] == pytest.approx(M.f1_at_threshold(yt, pt, thr), abs=2e-3)
        assert got["precision_at_k"] == pytest.approx(M.precision_at_k(yt, pt, 50), abs=0.021)
        assert got["recall_at_precision"] == pytest.approx(M.recall_at_precision(yt, pt, 0.90), abs=2e-3)
        assert got["ece"] == pytest.approx(M.expected_calibration_error(yt, pt), abs=1e-5)
        assert got["n_test"] == int(tm.sum()) and got["pr_auc_illicit"] > 0.15


def test_eval_step_graph_equals_eager(egnn, small_graph):
    """train.EvalStep: the eval_split forward as one CUDA graph reproduces the eager forward bit for bit and follows
    in-place changes of the parameters and the BatchNorm running statistics."""
    from egnn_b200.train import EvalStep, eval_probs
    gr = small_graph
    cfg = dict(hidden_dim=64, layers=3, dropout=0.2, time_embed_dim=2, time_embed_type="sin", max_timestep=49)
    torch.manual_seed(5)
    model = egnn.build_model("sage_resbn", 166, cfg).cuda()
    x, t = gr.x.cuda(), gr.timestep.cuda()
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
    ev = EvalStep(model, x, ei, t).capture()
    for it in range(3):
        p_e, l_e = eval_probs(model, x, ei, t)
        p_g, l_g = ev.run()
        assert torch.equal(l_g, l_e) and torch.equal(p_g, p_e), it
        with torch.no_grad():                  # what a training step does between two evaluations
            for p in model.parameters():
                p.mul_(1.01)
            for bn in model.bns:
                bn.running_mean.add_(0.05)
                bn.running_var.mul_(1.1)
    assert not model.training
