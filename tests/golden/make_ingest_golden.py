"""Golden vectors for the ingestion join, produced by the REFERENCE's own loader:
imports /root/reference/src/data/dataset_elliptic.py in the build container (`torch_geometric.data.Data`, which the
loader uses purely as a container, is replaced by an attribute bag because PyG is not installable here), writes seeded
CSV tables with the Elliptic layout into a temp dir, runs `load_elliptic_as_graph` + `make_temporal_masks` and records
inputs and outputs.  Run once here; the GPU box only reads tests/golden/ingest_golden.pt.
    python tests/golden/make_ingest_golden.py"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch


class Data:                                   # the container the loader fills (dataset_elliptic.py:247-248)
    def __init__(self, **kw):
        self.__dict__.update(kw)


tg, tgd = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.data")
tgd.Data = Data
tg.data = tgd
sys.modules["torch_geometric"], sys.modules["torch_geometric.data"] = tg, tgd
sys.path.insert(0, "/root/reference")
from src.data.dataset_elliptic import load_elliptic_as_graph, make_temporal_masks  # noqa: E402


def write_case(d, seed, n, f, e, n_t, *, ts_in, edge_header, unknown_frac, cross_frac, dup_tx, missing_cls):
    r = np.random.default_rng(seed)
    tx = r.choice(np.arange(1_000, 1_000 + 50 * n), size=n, replace=False).astype(np.int64) * 7919
    if dup_tx:                                 # a txId on two feature rows: the dict keeps the last
        tx[n - 1] = tx[1]
    t = np.sort(r.integers(1, n_t + 1, size=n)).astype(np.int64)
    x = r.standard_normal((n, f)).astype(np.float32)
    cls = r.choice(["1", "2", "unknown"], size=n, p=[0.1, 0.3, 0.6])
    # edges: mostly intra-timestep pairs, some crossing, some with endpoints outside the table, a self-loop, a repeat
    src = r.integers(0, n, size=e)
    dst = np.empty(e, dtype=np.int64)
    for i in range(e):
        same = np.flatnonzero(t == t[src[i]])
        dst[i] = r.choice(same)
    cross = r.random(e) < cross_frac
    dst[cross] = r.integers(0, n, size=int(cross.sum()))
    es, ed = tx[src].copy(), tx[dst].copy()
    unk = r.random(e) < unknown_frac
    es[unk] = 13                                # not a txId of the table
    if e > 4:
        es[3], ed[3] = es[2], ed[2]            # repeated edge stays repeated
        ed[4] = es[4]                          # self-loop
    with open(os.path.join(d, "elliptic_txs_features.csv"), "w") as fh:
        for i in range(n):
            cols = [str(tx[i])] + ([str(t[i])] if ts_in in ("features", "both") else []) + [repr(float(v)) for v in x[i]]
            fh.write(",".join(cols) + "\n")
    with open(os.path.join(d, "elliptic_txs_classes.csv"), "w") as fh:
        fh.write("txId,class" + (",time_step" if ts_in in ("classes", "both") else "") + "\n")
        for i in range(n):
            if missing_cls and i % 11 == 5:
                continue                       # node without a classes row -> label -1
            if dup_tx and i == n - 1:
                continue                       # one classes row per txId (a second one would duplicate feature rows)
            fh.write(f"{tx[i]},{cls[i]}" + (f",{t[i]}" if ts_in in ("classes", "both") else "") + "\n")
    with open(os.path.join(d, "elliptic_txs_edgelist.csv"), "w") as fh:
        if edge_header:
            fh.write("txId1,txId2\n")
        for a, b in zip(es, ed):
            fh.write(f"{a},{b}\n")
    return tx, es, ed


cases = []
specs = [
    dict(seed=0, n=40, f=5, e=60, n_t=5, ts_in="features", edge_header=True, unknown_frac=0.1, cross_frac=0.2, dup_tx=False, missing_cls=False),
    dict(seed=1, n=300, f=7, e=700, n_t=12, ts_in="features", edge_header=False, unknown_frac=0.05, cross_frac=0.1, dup_tx=True, missing_cls=True),
    dict(seed=2, n=2500, f=4, e=6000, n_t=49, ts_in="features", edge_header=True, unknown_frac=0.0, cross_frac=0.0, dup_tx=False, missing_cls=False),
    dict(seed=3, n=64, f=3, e=50, n_t=6, ts_in="features", edge_header=True, unknown_frac=1.0, cross_frac=0.0, dup_tx=False, missing_cls=False),
    dict(seed=4, n=500, f=6, e=0, n_t=9, ts_in="features", edge_header=True, unknown_frac=0.0, cross_frac=0.0, dup_tx=False, missing_cls=True),
]
for sp in specs:
    with tempfile.TemporaryDirectory() as d:
        tx, es, ed = write_case(d, **sp)
        with contextlib.redirect_stdout(io.StringIO()) as out:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                data, meta = load_elliptic_as_graph(d)
        log = out.getvalue()
        masks = {}
        for name, (a, b, k) in {"plain": (3, 4, None), "k2": (4, 5, 2), "k_big": (3, 4, 40)}.items():
            make_temporal_masks(data, a, b, k)
            masks[name] = {"args": (a, b, k), "train": data.train_mask.clone(), "val": data.val_mask.clone(),
                           "test": data.test_mask.clone()}
        files = {f: open(os.path.join(d, f)).read() for f in os.listdir(d)}
        cases.append({"spec": sp, "files": files, "tx_ids": torch.from_numpy(tx), "e_src_tx": torch.from_numpy(es),
                      "e_dst_tx": torch.from_numpy(ed), "x": data.x, "y": data.y, "timestep": data.timestep,
                      "edge_index": data.edge_index.reshape(2, -1), "meta": meta, "masks": masks, "log": log})
        print(sp["seed"], meta, log.strip().splitlines()[-1])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest_golden.pt")
torch.save(cases, out)
print("wrote", out, os.path.getsize(out), "bytes")
