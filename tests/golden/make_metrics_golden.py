"""Golden vectors for the epoch-tail metric, produced by the REFERENCE's own function:
imports /root/reference/src/utils/metrics.py (numpy + scikit-learn only) in the build container and records
`pr_auc_illicit` on seeded inputs.  Run once here; the GPU box only reads tests/golden/metrics_golden.json.
    python tests/golden/make_metrics_golden.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from src.utils.metrics import (expected_calibration_error, f1_at_threshold, pick_threshold_for_precision,  # noqa: E402
                               pick_threshold_max_f1, pr_auc_illicit, precision_at_k, recall_at_precision,
                               roc_auc_illicit)
from src.utils.calibrate import TemperatureScaler  # noqa: E402


def case(seed, n, pos_rate, quant):
    r = np.random.default_rng(seed)
    y = (r.random(n) < pos_rate).astype(np.int64)
    s = (r.random(n) * 0.6 + 0.4 * y * r.random(n)).astype(np.float32)
    if quant:
        s = (np.round(s * quant) / quant).astype(np.float32)   # heavy ties
    return y, s


cases = []
for seed, n, pr, q in [(0, 8, 0.4, 0), (1, 50, 0.1, 4), (2, 1000, 0.02, 0), (3, 1000, 0.1, 16), (4, 20000, 0.1, 256),
                       (5, 20000, 0.022, 0), (6, 257, 0.5, 2), (7, 4097, 0.3, 0)]:
    y, s = case(seed, n, pr, q)
    thr, f1 = pick_threshold_max_f1(y, s)
    k = min(100, max(1, n // 3))
    cases.append({"seed": seed, "n": n, "pos_rate": pr, "quant": q, "ap": pr_auc_illicit(y, s),
                  "roc": roc_auc_illicit(y, s), "thr_max_f1": thr, "max_f1": f1,
                  "thr_p50": pick_threshold_for_precision(y, s, 0.5), "f1_at_thr": f1_at_threshold(y, s, thr),
                  "f1_at_045": f1_at_threshold(y, s, 0.45), "k": k,
                  # precision_at_k is tie-order dependent in the reference (unstable argsort): golden only without ties
                  "p_at_k": precision_at_k(y, s, k) if not q else None,
                  "rec_at_p50": recall_at_precision(y, s, 0.5), "rec_at_p90": recall_at_precision(y, s, 0.9),
                  "ece": expected_calibration_error(y, s)})
# the reference's own test vector (tests/test_masks_and_metrics.py:22-25)
y = np.array([0, 1, 0, 1, 0, 0, 0, 1])
s = np.linspace(0, 1, len(y))
cases.append({"reference_test": True, "y": y.tolist(), "s": s.tolist(), "ap": pr_auc_illicit(y, s),
              "roc": roc_auc_illicit(y, s)})
# temperature scaling: the reference's own TemperatureScaler.fit (LBFGS on T) on seeded over-/under-confident logits
import torch  # noqa: E402

temps = []
# (2, 9000, 1.0): the reference's LBFGS (lr 0.1 from T = 1) overshoots into T < 0 and returns T = -2552 with an NLL of
# 0.693 (worse than uncalibrated); recorded as `reference_diverged` -- the device solver returns the minimiser there
for seed, n, scale in [(0, 4000, 3.0), (1, 4000, 0.4), (2, 9000, 1.0), (2, 9000, 1.5), (3, 500, 6.0), (4, 9000, 10.0)]:
    g = torch.Generator().manual_seed(seed)
    yy = (torch.rand(n, generator=g) < 0.15).long()
    margin = torch.randn(n, generator=g) + 1.2 * (2 * yy.float() - 1)      # informative but noisy
    logits = torch.stack([-0.5 * margin, 0.5 * margin], dim=1) * scale
    torch.manual_seed(0)
    T = TemperatureScaler().fit(logits, yy)
    temps.append({"seed": seed, "n": n, "scale": scale, "T": T, "reference_diverged": bool(T <= 0)})

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_golden.json")
json.dump({"generator": "tests/golden/make_metrics_golden.py",
           "function": "src.utils.metrics.* / src.utils.calibrate.TemperatureScaler.fit",
           "cases": cases, "temperature": temps}, open(out, "w"), indent=1)
print("wrote", out, [round(c["ap"], 6) for c in cases])
