"""Golden vectors for the epoch-tail metric, produced by the REFERENCE's own function:
imports /root/reference/src/utils/metrics.py (numpy + scikit-learn only) in the build container and records
`pr_auc_illicit` on seeded inputs.  Run once here; the GPU box only reads tests/golden/metrics_golden.json.
    python tests/golden/make_metrics_golden.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from src.utils.metrics import pr_auc_illicit, roc_auc_illicit  # noqa: E402


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
    cases.append({"seed": seed, "n": n, "pos_rate": pr, "quant": q, "ap": pr_auc_illicit(y, s),
                  "roc": roc_auc_illicit(y, s)})
# the reference's own test vector (tests/test_masks_and_metrics.py:22-25)
y = np.array([0, 1, 0, 1, 0, 0, 0, 1])
s = np.linspace(0, 1, len(y))
cases.append({"reference_test": True, "y": y.tolist(), "s": s.tolist(), "ap": pr_auc_illicit(y, s),
              "roc": roc_auc_illicit(y, s)})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_golden.json")
json.dump({"generator": "tests/golden/make_metrics_golden.py",
           "function": "src.utils.metrics.pr_auc_illicit / roc_auc_illicit",
           "cases": cases}, open(out, "w"), indent=1)
print("wrote", out, [round(c["ap"], 6) for c in cases])
