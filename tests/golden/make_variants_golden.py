"""Generate tests/golden/variants_golden.pt by running the REFERENCE's own graph-variant code, unmodified, imported
from /root/reference:
  * src.analysis.hub_ablation.build_edge_index_ablated   (hub_ablation.py:56-71)
  * src.analysis.robustness.drop_edges                   (robustness.py:65-82)
Both modules import torch_geometric (absent here) only for the `Data` type and the model zoo, so stub modules are
registered first, as tests/golden/make_golden.py does; the two functions themselves are pure torch.

Run in the build container only:  python tests/golden/make_variants_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden as MG  # noqa: E402  (stubs + sys.path entry for /root/reference)

MG._stub()
from src.analysis.hub_ablation import build_edge_index_ablated  # noqa: E402
from src.analysis.robustness import drop_edges  # noqa: E402

sys.path.insert(0, os.path.join(ROOT))
import egnn_b200  # noqa: E402,F401  (import shim for the hyphenated package directory)
from egnn_b200 import synthetic  # noqa: E402


def graphs():
    adv = synthetic.adversarial_tiny()
    small = synthetic.make_elliptic_like(n_nodes=500, n_edges=600, n_feats=4, n_timesteps=6, seed=7, hub_degree=40,
                                         t_train_end=4, t_val_end=5)
    path = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]])      # the reference's own fixture (tests/test_masks_and_metrics.py:12)
    sym = lambda ei: torch.cat([ei, ei.flip(0)], 1)
    return {"path": (5, path), "adversarial": (adv.num_nodes, adv.edge_index),
            "small": (small.num_nodes, small.edge_index), "small_sym": (small.num_nodes, sym(small.edge_index))}


def main():
    out = {"graphs": {}, "ablate": [], "drop": []}
    for name, (n, ei) in graphs().items():
        out["graphs"][name] = {"num_nodes": n, "edge_index": ei.to(torch.int32)}
        for frac in (0.0, 0.001, 0.01, 0.05, 0.2, 0.5, 1.0):
            abl, k = build_edge_index_ablated(ei, n, frac, torch.device("cpu"))
            out["ablate"].append({"graph": name, "frac": frac, "num_hubs": k, "edge_index": abl.to(torch.int32)})
        for frac in (0.0, 1e-4, 0.1, 0.125, 0.5, 0.9):
            for seed in (0, 5):
                torch.manual_seed(seed)
                try:
                    kept, cnt = drop_edges(ei, frac)
                    out["drop"].append({"graph": name, "frac": frac, "seed": seed, "count": cnt,
                                        "edge_index": kept.to(torch.int32)})
                except RuntimeError as e:
                    out["drop"].append({"graph": name, "frac": frac, "seed": seed, "error": str(e)})
    # node ids < 2^31: stored as int32 to keep the fixture small (the tests widen them back)
    path = os.path.join(HERE, "variants_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["ablate"]), "ablations,", len(out["drop"]), "drops")


if __name__ == "__main__":
    main()
