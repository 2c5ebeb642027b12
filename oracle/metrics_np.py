"""NumPy oracle for the epoch-tail metric  --  TEST INFRASTRUCTURE ONLY.

Restates `pr_auc_illicit` (`/root/reference/src/utils/metrics.py:11-13`), i.e. scikit-learn's
`average_precision_score` for a binary target (third-party, unpinned in the reference's `environment.yml:9`;
algorithm as published in `sklearn/metrics/_ranking.py`: `_binary_clf_curve` -> `precision_recall_curve` ->
`-sum(diff(recall) * precision[:-1])`).  PARITY PINNED: checked in tests/test_oracle_metrics.py against
(a) scikit-learn itself (installed here and on the GPU box) and (b) golden vectors produced by importing the
reference's own `src/utils/metrics.py` in this container (`tests/golden/make_metrics_golden.py`).

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
