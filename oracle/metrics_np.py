"""NumPy oracle for the epoch-tail and final metrics  --  TEST INFRASTRUCTURE ONLY.

Restates `pr_auc_illicit` (`/root/reference/src/utils/metrics.py:11-13`), i.e. scikit-learn's
`average_precision_score` for a binary target (third-party, unpinned in the reference's `environment.yml:9`;
algorithm as published in `sklearn/metrics/_ranking.py`: `_binary_clf_curve` -> `precision_recall_curve` ->
`-sum(diff(recall) * precision[:-1])`).  PARITY PINNED: checked in tests/test_oracle_metrics.py against
(a) scikit-learn itself (installed here and on the GPU box) and (b) golden vectors produced by importing the
reference's own `src/utils/metrics.py` in this container (`tests/golden/make_metrics_golden.py`).

The final-metrics functions (`pick_threshold_max_f1`, `pick_threshold_for_precision`, `f1_at_threshold`,
`precision_at_k`, `recall_at_precision`, `expected_calibration_error`, `/root/reference/src/utils/metrics.py:18-66`) and
the temperature-scaling objective (`/root/reference/src/utils/calibrate.py:8-30`) are pinned the same way: golden vectors
from the reference's own functions / `TemperatureScaler.fit` (`tests/golden/make_metrics_golden.py`).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def binary_clf_curve(y_true: np.ndarray, y_score: np.ndarray):
    """fps, tps, thresholds at the distinct score values, descending (sklearn `_binary_clf_curve`)."""
    y_true = np.asarray(y_true).astype(np.int64)
    y_score = np.asarray(y_score)
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_true = y_score[order], y_true[order]
    distinct = np.where(np.diff(y_score))[0]
    idx = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true, dtype=np.float64)[idx]
    fps = 1 + idx - tps
    return fps, tps, y_score[idx]


def average_precision(y_true: np.ndarray, y_score: np.ndarray):
    """-> (AP, n, positives, distinct thresholds).  No sample or no positive: 0.0 (the reference guards the empty
    case itself, `src/train_gnn.py:390-392`; sklearn defines recall = 1 everywhere without positives -> 0.0)."""
    y_true = np.asarray(y_true)
    if y_true.size == 0:
        return 0.0, 0, 0, 0
    fps, tps, thr = binary_clf_curve(y_true, y_score)
    n_pos = int(tps[-1])
    if n_pos == 0:
        return 0.0, int(y_true.size), 0, int(thr.size)
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    precision = np.hstack((precision[::-1], 1.0))
    recall = np.hstack((recall[::-1], 0.0))
    ap = float(max(0.0, -np.sum(np.diff(recall) * precision[:-1])))
    return ap, int(y_true.size), n_pos, int(thr.size)


def roc_auc(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """`roc_auc_illicit` (`/root/reference/src/utils/metrics.py:15-16`) = sklearn `roc_auc_score` for a binary target:
    trapezoid area under (fpr, tpr) over the distinct thresholds, starting at (0, 0).  NaN when only one class is
    present (sklearn raises ValueError there)."""
    y_true = np.asarray(y_true)
    if y_true.size == 0:
        return float("nan")
    fps, tps, _ = binary_clf_curve(y_true, y_score)
    if tps[-1] == 0 or fps[-1] == 0:
        return float("nan")
    fpr = np.r_[0.0, fps / fps[-1]]
    tpr = np.r_[0.0, tps / tps[-1]]
    return float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) * 0.5))


def precision_recall_curve(y_true: np.ndarray, y_score: np.ndarray):
    """sklearn `precision_recall_curve` (no `drop_intermediate`): precision, recall over ASCENDING thresholds with the
    final (1, 0) point appended; recall is 1 everywhere when there is no positive."""
    fps, tps, thr = binary_clf_curve(y_true, y_score)
    ps = tps + fps
    precision = np.zeros_like(tps)
    np.divide(tps, ps, out=precision, where=(ps != 0))
    recall = np.ones_like(tps) if tps[-1] == 0 else tps / tps[-1]
    return np.hstack((precision[::-1], 1.0)), np.hstack((recall[::-1], 0.0)), thr[::-1]


def pick_threshold_max_f1(y_true: np.ndarray, y_score: np.ndarray):
    """`/root/reference/src/utils/metrics.py:22-27` -> (threshold, f1)."""
    precision, recall, thresholds = precision_recall_curve(y_true, y_score)
    thresholds = np.append(thresholds, 1.0)
    f1s = 2 * precision * recall / (precision + recall + 1e-12)
    i = int(np.nanargmax(f1s))
    return float(thresholds[i]), float(f1s[i])


def pick_threshold_for_precision(y_true: np.ndarray, y_score: np.ndarray, target_p: float) -> float:
    """`/root/reference/src/utils/metrics.py:29-37`."""
    precision, recall, thresholds = precision_recall_curve(y_true, y_score)
    cand = np.append(thresholds, 1.0)
    mask = precision >= target_p
    if not np.any(mask):
        return pick_threshold_max_f1(y_true, y_score)[0]
    return float(cand[int(np.argmax(mask))])


def f1_at_threshold(y_true: np.ndarray, y_score: np.ndarray, thr: float) -> float:
    """`/root/reference/src/utils/metrics.py:18-20` (sklearn `f1_score`, binary): 2 tp / (2 tp + fp + fn), 0 if undefined."""
    y_true = np.asarray(y_true).astype(np.int64)
    pred = (np.asarray(y_score) >= thr).astype(np.int64)
    tp = int((pred & y_true).sum())
    den = int(pred.sum()) + int(y_true.sum())
    return 2.0 * tp / den if den > 0 else 0.0


def precision_at_k(y_true: np.ndarray, y_score: np.ndarray, k: int) -> float:
    """`/root/reference/src/utils/metrics.py:39-41` (ties at the k-th score are broken by a STABLE descending sort
    here; the reference's `np.argsort(-s)` leaves them unspecified)."""
    idx = np.argsort(-np.asarray(y_score), kind="stable")[:k]
    return float(np.mean(np.asarray(y_true)[idx]))


def recall_at_precision(y_true: np.ndarray, y_score: np.ndarray, target_p: float) -> float:
    """`/root/reference/src/utils/metrics.py:43-48`."""
    precision, recall, _ = precision_recall_curve(y_true, y_score)
    mask = precision >= target_p
    return float(np.max(recall[mask])) if np.any(mask) else 0.0


def expected_calibration_error(y_true: np.ndarray, y_prob: np.ndarray, bins: int = 15) -> float:
    """`/root/reference/src/utils/metrics.py:50-66`."""
    y_true = np.asarray(y_true).astype(int)
    y_prob = np.asarray(y_prob)
    edges = np.linspace(0.0, 1.0, bins + 1)
    ece = 0.0
    for i in range(bins):
        lo, hi = edges[i], edges[i + 1]
        m = (y_prob >= lo) & ((y_prob < hi) if i < bins - 1 else (y_prob <= hi))
        if not np.any(m):
            continue
        ece += m.mean() * abs(y_true[m].mean() - y_prob[m].mean())
    return float(ece)


def temperature_nll(logits: np.ndarray, y: np.ndarray, T: float) -> float:
    """mean CrossEntropyLoss(logits / T, y) in float64 (the objective of `src/utils/calibrate.py:17-27`)."""
    z = np.asarray(logits, dtype=np.float64) / T
    m = z.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(z - m).sum(axis=1, keepdims=True)))[:, 0]
    return float(np.mean(lse - z[np.arange(len(y)), np.asarray(y).astype(int)]))


def fit_temperature(logits: np.ndarray, y: np.ndarray) -> float:
    """The minimiser the reference's LBFGS converges to: argmin_T temperature_nll(logits, y, T), by golden-section
    search on beta = 1/T (the objective is convex in beta)."""
    lo, hi = 1e-4, 1e3
    phi = (np.sqrt(5.0) - 1.0) / 2.0
    f = lambda b: temperature_nll(logits, y, 1.0 / b)
    a, b = lo, hi
    c, d = b - phi * (b - a), a + phi * (b - a)
    fc, fd = f(c), f(d)
    for _ in range(200):
        if fc < fd:
            b, d, fd = d, c, fc
            c = b - phi * (b - a)
            fc = f(c)
        else:
            a, c, fc = c, d, fd
            d = a + phi * (b - a)
            fd = f(d)
    return 2.0 / (a + b)


def softmax_pos(logits: np.ndarray) -> np.ndarray:
    """softmax(logits, 1)[:, 1] in float32 (`eval_split`, `src/train_gnn.py:254`)."""
    l = np.asarray(logits, dtype=np.float32)
    m = l.max(axis=1, keepdims=True)
    e = np.exp(l - m, dtype=np.float32)
    return (e[:, 1] / e.sum(axis=1, dtype=np.float32)).astype(np.float32)


class EarlyStop:
    """best_val / bad / best epoch of `src/train_gnn.py:375-411`."""

    def __init__(self):
        self.best, self.bad, self.best_epoch, self.epoch = -1.0, 0, 0, 0

    def update(self, v: float) -> bool:
        self.epoch += 1
        if v > self.best:
            self.best, self.bad, self.best_epoch = v, 0, self.epoch
            return True
        self.bad += 1
        return False
