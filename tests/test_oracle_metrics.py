"""The epoch-tail oracle (oracle/metrics_np.py) pinned against the reference's own function (golden vectors made by
importing /root/reference/src/utils/metrics.py, tests/golden/make_metrics_golden.py) and against scikit-learn."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import metrics_np as M

HERE = os.path.dirname(os.path.abspath(__file__))


def _gen():
    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_metrics_golden.py"))
    return spec


def _case(seed, n, pos_rate, quant):
    r = np.random.default_rng(seed)
    y = (r.random(n) < pos_rate).astype(np.int64)
    s = (r.random(n) * 0.6 + 0.4 * y * r.random(n)).astype(np.float32)
    if quant:
        s = (np.round(s * quant) / quant).astype(np.float32)
    return y, s


def test_golden_vectors_from_the_reference_function():
    g = json.load(open(os.path.join(HERE, "golden", "metrics_golden.json")))
    assert len(g["cases"]) >= 9
    for c in g["cases"]:
        if c.get("reference_test"):
            y, s = np.array(c["y"]), np.array(c["s"])
        else:
            y, s = _case(c["seed"], c["n"], c["pos_rate"], c["quant"])
        ap = M.average_precision(y, s)[0]
        assert ap == pytest.approx(c["ap"], rel=1e-13, abs=1e-15), c
        assert M.roc_auc(y, s) == pytest.approx(c["roc"], rel=1e-13), c


def test_final_metrics_golden_from_the_reference_functions():
    """pick_threshold_max_f1 / pick_threshold_for_precision / f1_at_threshold / precision_at_k / recall_at_precision /
    expected_calibration_error (`/root/reference/src/utils/metrics.py:18-66`), recorded from the reference itself."""
    g = json.load(open(os.path.join(HERE, "golden", "metrics_golden.json")))
    n_checked = 0
    for c in g["cases"]:
        if c.get("reference_test"):
            continue
        y, s = _case(c["seed"], c["n"], c["pos_rate"], c["quant"])
        thr, f1 = M.pick_threshold_max_f1(y, s)
        assert thr == c["thr_max_f1"] and f1 == pytest.approx(c["max_f1"], rel=1e-13), c
        assert M.pick_threshold_for_precision(y, s, 0.5) == c["thr_p50"], c
        assert M.f1_at_threshold(y, s, thr) == pytest.approx(c["f1_at_thr"], rel=1e-13), c
        assert M.f1_at_threshold(y, s, 0.45) == pytest.approx(c["f1_at_045"], rel=1e-13), c
        if c["p_at_k"] is not None:
            assert M.precision_at_k(y, s, c["k"]) == pytest.approx(c["p_at_k"], rel=1e-13), c
        assert M.recall_at_precision(y, s, 0.5) == pytest.approx(c["rec_at_p50"], rel=1e-13), c
        assert M.recall_at_precision(y, s, 0.9) == pytest.approx(c["rec_at_p90"], rel=1e-13), c
        assert M.expected_calibration_error(y, s) == pytest.approx(c["ece"], rel=1e-12), c
        n_checked += 1
    assert n_checked >= 8


def _temp_case(seed, n, scale):
    import torch
    g = torch.Generator().manual_seed(seed)
    yy = (torch.rand(n, generator=g) < 0.15).long()
    margin = torch.randn(n, generator=g) + 1.2 * (2 * yy.float() - 1)
    logits = torch.stack([-0.5 * margin, 0.5 * margin], dim=1) * scale
    return logits, yy


def test_temperature_oracle_against_the_reference_scaler():
    """`TemperatureScaler.fit` (`/root/reference/src/utils/calibrate.py:8-30`) is LBFGS in fp32 and stops within ~1e-3
    of the minimiser; the oracle's minimiser must reach an NLL at least as low and a T within 2e-3 relative.  Where
    the reference's LBFGS diverges (T < 0, recorded) only the objective is compared."""
    g = json.load(open(os.path.join(HERE, "golden", "metrics_golden.json")))
    assert len(g["temperature"]) >= 5
    for c in g["temperature"]:
        logits, yy = _temp_case(c["seed"], c["n"], c["scale"])
        T = M.fit_temperature(logits.numpy(), yy.numpy())
        nll = M.temperature_nll(logits.numpy(), yy.numpy(), T)
        assert nll <= M.temperature_nll(logits.numpy(), yy.numpy(), c["T"]) + 1e-12
        if not c["reference_diverged"]:
            assert T == pytest.approx(c["T"], rel=2e-3), c


def test_against_sklearn_including_ties_and_degenerate_inputs():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(11)
    for n, q in [(1, 0), (2, 0), (17, 2), (300, 8), (5000, 0), (5000, 64)]:
        y = (rng.random(n) < 0.3).astype(np.int64)
        if y.sum() == 0:
            y[0] = 1
        s = rng.random(n).astype(np.float32)
        if q:
            s = (np.round(s * q) / q).astype(np.float32)
        ap, cnt, pos, thr = M.average_precision(y, s)
        assert ap == pytest.approx(float(sk.average_precision_score(y, s)), rel=1e-13)
        assert (cnt, pos, thr) == (n, int(y.sum()), np.unique(s).size)
        if 0 < y.sum() < n:
            assert M.roc_auc(y, s) == pytest.approx(float(sk.roc_auc_score(y, s)), rel=1e-13)
        else:
            assert np.isnan(M.roc_auc(y, s))
    # all scores equal: one threshold, AP = prevalence
    y = np.array([1, 0, 0, 1, 0])
    assert M.average_precision(y, np.full(5, 0.5, np.float32))[0] == pytest.approx(0.4)
    # no positives / empty selection -> 0.0 (sklearn warns and returns 0.0; the reference guards size == 0)
    assert M.average_precision(np.zeros(4, np.int64), np.arange(4, dtype=np.float32))[0] == 0.0
    assert M.average_precision(np.zeros(0, np.int64), np.zeros(0, np.float32)) == (0.0, 0, 0, 0)


def test_early_stop_bookkeeping():
    es = M.EarlyStop()
    seq = [0.2, 0.3, 0.3, 0.25, 0.31, 0.1]
    flags = [es.update(v) for v in seq]
    assert flags == [True, True, False, False, True, False]          # strict `>` (src/train_gnn.py:394)
    assert (es.best, es.bad, es.best_epoch, es.epoch) == (0.31, 1, 5, 6)
