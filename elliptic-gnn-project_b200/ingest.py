"""The data formats in front of the hot path: the Elliptic CSV tables and the cached
`graph.pt` become device-resident `(x, edge_index, y, timestep, masks)` and the sorted CSR/CSC views directly.

Mirrors the reference interface (same function names, arguments, return convention and error behaviour):

* `load_elliptic_as_graph(data_dir, features_csv, classes_csv, edgelist_csv) -> (data, meta)`
  (`/root/reference/src/data/dataset_elliptic.py:49-265`).  The text work (pandas CSV parsing, the column-name
  heuristics, the label map) stays on the host exactly as the reference does it; the part the reference runs as a
  per-edge Python dictionary walk -- txId -> row index for both endpoints, dropping edges with an unknown endpoint
  or endpoints in different timesteps (`:190-245`) -- is ONE device pass (`egnn_txid_join`: hash build, probe,
  stable compaction), bit-exact with the reference's `edge_index`.
* `make_temporal_masks(data, t_train_end, t_val_end, train_window_k)` (`:268-290`) -> `egnn_temporal_masks`.
* `load_cached(processed_dir)` (`/root/reference/src/train_gnn.py:50-64`): reads `graph.pt`.  The reference pickles a
  `torch_geometric.data.Data`; PyG is not importable in this image, so the unpickler maps the PyG container classes
  to plain attribute bags and lifts the tensors out (also accepts a plain dict of tensors or an `EllipticGraph`).
* `to_device_graph(data, symmetrize_edges, self_loops)`: the tensors on the GPU + the sorted views registered in the
  graph cache, so that the first `model(x, edge_index, t)` call finds its CSR/CSC ready.

CUDA only: the join and the masks have no CPU path (the oracle in `oracle/ingest_np.py` is test infrastructure).
"""
from __future__ import annotations

import io
import os
import pickle
import warnings
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from ._lib import check, lib, ptr, stream
from .synthetic import EllipticGraph

# `LABEL_MAPS` / `_map_label` (dataset_elliptic.py:12-28): class1/1/illicit -> 1, class2/2/licit -> 0, else -1
LABEL_MAPS = {"class1": 1, "1": 1, 1: 1, "illicit": 1, "class2": 0, "2": 0, 2: 0, "licit": 0, "unknown": -1, -1: -1}


def _map_label(v):
    s = str(v).strip().lower()
    return LABEL_MAPS.get(s, LABEL_MAPS.get(v, -1))


def _looks_like_timestep(col) -> bool:
    """dataset_elliptic.py:31-46: integers within [1, 49] in > 95 % of the rows."""
    import pandas as pd
    if not np.issubdtype(col.dtype, np.number):
        try:
            col = pd.to_numeric(col, errors="coerce")
        except Exception:
            return False
    vals = col.dropna().astype(float)
    if vals.empty:
        return False
    return bool((vals.min() >= 1) and (vals.max() <= 49) and (vals.round().eq(vals).mean() > 0.95))


def join_edges(tx_ids: torch.Tensor, timestep: torch.Tensor, e_src_tx: torch.Tensor, e_dst_tx: torch.Tensor):
    """txId pairs -> node-index `edge_index` int64 [2, E_kept] on the device (dataset_elliptic.py:190-245).
    Returns (edge_index, {"mapped": both endpoints known, "kept": also same timestep, "duplicate_txids": extra rows})."""
    for t in (tx_ids, timestep, e_src_tx, e_dst_tx):
        if t.dtype != torch.int64 or t.dim() != 1:
            raise TypeError("join_edges takes 1-D int64 tensors")
        if not t.is_cuda:
            raise RuntimeError("egnn_b200 joins the edge list on the GPU only (no CPU fallback)")
    if timestep.numel() != tx_ids.numel() or e_src_tx.numel() != e_dst_tx.numel():
        raise ValueError("length mismatch")
    N, E = int(tx_ids.numel()), int(e_src_tx.numel())
    dev = tx_ids.device
    L = lib()
    out = torch.empty((2, max(E, 1)), dtype=torch.int64, device=dev)
    info = torch.empty(3, dtype=torch.int32, device=dev)
    nws = L.egnn_txid_join_workspace_bytes(N, E)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    check(L.egnn_txid_join(ptr(tx_ids.contiguous()), ptr(timestep.contiguous()), N, ptr(e_src_tx.contiguous()),
                           ptr(e_dst_tx.contiguous()), E, ptr(out), ptr(info), ptr(ws), nws, stream()))
    kept, mapped, dup = (int(v) for v in info.tolist())
    ei = out[:, :kept].contiguous() if E > 0 else out[:, :0]
    return ei, {"mapped": mapped, "kept": kept, "duplicate_txids": dup}


def make_temporal_masks(data, t_train_end: int, t_val_end: int, train_window_k: Optional[int] = None):
    """`make_temporal_masks` (dataset_elliptic.py:268-290) on the device; sets `data.train_mask / val_mask /
    test_mask` (bool [N]) and returns `data`, like the reference."""
    y, t = data.y, data.timestep
    if not (y.is_cuda and t.is_cuda):
        raise RuntimeError("egnn_b200 builds the masks on the GPU only (no CPU fallback)")
    n = int(y.numel())
    m = torch.empty((3, n), dtype=torch.uint8, device=y.device)
    check(lib().egnn_temporal_masks(ptr(y.contiguous()), ptr(t.contiguous()), n, int(t_train_end), int(t_val_end),
                                    -1 if train_window_k is None else int(train_window_k), ptr(m[0]), ptr(m[1]),
                                    ptr(m[2]), stream()))
    data.train_mask, data.val_mask, data.test_mask = m[0].bool(), m[1].bool(), m[2].bool()
    return data


def load_elliptic_as_graph(data_dir: str, features_csv: str = "elliptic_txs_features.csv",
                           classes_csv: str = "elliptic_txs_classes.csv",
                           edgelist_csv: str = "elliptic_txs_edgelist.csv",
                           device="cuda") -> Tuple[EllipticGraph, Dict]:
    """Same contract as the reference loader: data.x [N, F] float32, data.y [N] in {0, 1, -1}, data.edge_index
    [2, E] int64 (intra-timestep, CSV order), data.timestep [N]; node i = row i of the features CSV.  All four tensors
    live on `device`."""
    import pandas as pd
    f_path, c_path, e_path = (os.path.join(data_dir, f) for f in (features_csv, classes_csv, edgelist_csv))

    # ---- classes: txId + class (+ optional time_step / timestep)            (dataset_elliptic.py:69-106)
    df_cls = pd.read_csv(c_path)
    df_cls.columns = [c.strip() for c in df_cls.columns]
    if "txId" not in df_cls.columns:
        for col in df_cls.columns:
            if col.lower().startswith("tx"):
                df_cls = df_cls.rename(columns={col: "txId"})
                break
    if "time_step" in df_cls.columns:
        df_cls = df_cls.rename(columns={"time_step": "timestep"})
    has_cls_ts = "timestep" in df_cls.columns
    if "class" not in df_cls.columns:
        for col in df_cls.columns:
            if col.lower().startswith("class"):
                df_cls = df_cls.rename(columns={col: "class"})
                break
    df_cls["txId"] = pd.to_numeric(df_cls["txId"], errors="raise").astype(np.int64)
    if has_cls_ts:
        df_cls["timestep"] = pd.to_numeric(df_cls["timestep"], errors="raise").astype(np.int64)
    df_cls["label"] = df_cls["class"].apply(_map_label)
    df_cls = df_cls[["txId", "label"] + (["timestep"] if has_cls_ts else [])]

    # ---- features: headerless, txId | [timestep] | F features               (:111-151)
    df_feat = pd.read_csv(f_path, header=None)
    if df_feat.shape[1] < 2:
        raise ValueError("features CSV appears malformed (needs at least txId + 1 column).")
    tx_col = pd.to_numeric(df_feat.iloc[:, 0], errors="raise").astype(np.int64)
    feat_has_ts = _looks_like_timestep(df_feat.iloc[:, 1])
    feat_timestep = pd.to_numeric(df_feat.iloc[:, 1], errors="raise").astype(np.int64) if feat_has_ts else None
    feats = df_feat.iloc[:, (2 if feat_has_ts else 1):]
    if not has_cls_ts and not feat_has_ts:
        raise ValueError(
            "No timestep column found in classes and features did not contain a valid timestep column.\n"
            "Expected either classes.csv to have 'time_step'/'timestep' OR features.csv column 2 to be 1..49.")

    # ---- left join features <- classes on txId                              (:156-181)
    left = pd.DataFrame({"txId": tx_col.values, "_row": np.arange(len(tx_col), dtype=np.int64)})
    if not has_cls_ts:
        left["timestep"] = feat_timestep.values   # classes carry no time column: take the features' (:144-146)
    used_ts_source = "CLASSES" if has_cls_ts else "FEATURES"
    df = left.merge(df_cls, on="txId", how="left")  # a txId listed twice in classes.csv duplicates its row, as there
    if "timestep" not in df.columns:
        raise ValueError("Failed to construct 'timestep' from available CSVs.")
    ts = df["timestep"]
    feats = feats.iloc[df["_row"].values]
    df["label"] = df["label"].fillna(-1).astype(int)
    print(f"[TS] using timestep from: {used_ts_source}")

    # ---- tensors on the device                                             (:186-196)
    x = torch.tensor(feats.values, dtype=torch.float32).to(device)
    y = torch.tensor(df["label"].values, dtype=torch.int64).to(device)
    timestep = torch.tensor(ts.values.astype(np.int64), dtype=torch.int64).to(device)
    tx_ids = torch.from_numpy(df["txId"].values.astype(np.int64)).to(device)

    # ---- edge list: header 'txId1,txId2' or none                             (:198-219)
    try:
        sniff = pd.read_csv(e_path, nrows=5)
        if sniff.shape[1] >= 2 and not np.issubdtype(sniff.dtypes.iloc[0], np.number):
            df_edge = pd.read_csv(e_path, header=0)
        else:
            df_edge = pd.read_csv(e_path, header=None)
    except Exception:
        df_edge = pd.read_csv(e_path, header=None)
    if {"txId1", "txId2"}.issubset(set(df_edge.columns)):
        df_edge = df_edge[["txId1", "txId2"]].copy()
    else:
        df_edge = df_edge.iloc[:, :2].copy()
    df_edge.columns = ["src", "dst"]
    df_edge["src"] = pd.to_numeric(df_edge["src"], errors="coerce").astype("Int64")
    df_edge["dst"] = pd.to_numeric(df_edge["dst"], errors="coerce").astype("Int64")
    df_edge = df_edge.dropna().astype({"src": "int64", "dst": "int64"})
    edges_total = len(df_edge)

    # ---- txId -> row index, intra-timestep filter: on the device             (:221-245)
    e_src = torch.from_numpy(df_edge["src"].values.astype(np.int64)).to(device)
    e_dst = torch.from_numpy(df_edge["dst"].values.astype(np.int64)).to(device)
    edge_index, counts = join_edges(tx_ids, timestep, e_src, e_dst)
    if counts["mapped"] == 0:
        warnings.warn("No edges mapped to known txIds. If you are testing with a small/partial features CSV, "
                      "this is expected. Use the full features file to see edges.")
    print(f"[EDGES] total_in_csv={edges_total} mapped={counts['mapped']} same_t={counts['kept']} "
          f"kept_in_graph={edge_index.size(1)}")

    data = EllipticGraph(x=x, edge_index=edge_index, y=y, timestep=timestep)
    meta = {
        "num_nodes": int(x.size(0)),
        "num_edges": int(edge_index.size(1)),
        "num_features": int(x.size(1)),
        "label_counts": {"-1": int((y == -1).sum()), "0": int((y == 0).sum()), "1": int((y == 1).sum())},
    }
    return data, meta


# ------------------------------------------------------------------------------------------ graph.pt
class _Bag:
    """Stand-in for the PyG container classes inside a pickled `Data` (Data, GlobalStorage, ...): keeps whatever
    state the pickle sets, so the tensors can be lifted out without importing torch_geometric."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["_state"] = state
        if isinstance(state, dict):
            self.__dict__.update(state)


class _PygFreeUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "torch_geometric":
            return type(name, (_Bag,), {})
        return super().find_class(module, name)


class _pickle_shim:
    """`pickle_module` for torch.load: the stock pickle with PyG classes mapped to attribute bags."""
    __name__ = "egnn_b200_pickle_shim"
    Unpickler = _PygFreeUnpickler
    load = staticmethod(lambda f, **k: _PygFreeUnpickler(f, **k).load())
    loads = staticmethod(lambda b, **k: _PygFreeUnpickler(io.BytesIO(b), **k).load())
    dumps, dump, PickleError, UnpicklingError = pickle.dumps, pickle.dump, pickle.PickleError, pickle.UnpicklingError


_FIELDS = ("x", "edge_index", "y", "timestep", "train_mask", "val_mask", "test_mask")


def _lift(obj) -> Dict[str, torch.Tensor]:
    """Find the tensor fields of a loaded object: a dict, an object with the attributes, or a PyG-style bag whose
    storage keeps them in `_store._mapping`."""
    seen, stack = set(), [obj]
    while stack:
        o = stack.pop()
        if id(o) in seen:
            continue
        seen.add(id(o))
        d = o if isinstance(o, dict) else getattr(o, "__dict__", None)
        if not isinstance(d, dict):
            continue
        if isinstance(d.get("x"), torch.Tensor) and isinstance(d.get("edge_index"), torch.Tensor):
            return {k: d[k] for k in _FIELDS if isinstance(d.get(k), torch.Tensor)}
        stack.extend(v for v in d.values() if isinstance(v, (dict, _Bag)) or hasattr(v, "__dict__"))
    raise RuntimeError("no (x, edge_index) tensors found in the loaded object")


def load_cached(processed_dir: str) -> EllipticGraph:
    """`load_cached` (src/train_gnn.py:50-64): `processed_dir/graph.pt` -> host tensors (map_location cpu, like the
    reference; `to_device_graph` moves them).  Same error text on failure."""
    path = os.path.join(processed_dir, "graph.pt")
    try:
        try:
            import torch_geometric  # noqa: F401  (when present the stock unpickler yields a real Data object)
            obj = torch.load(path, map_location="cpu", weights_only=False)
        except ImportError:
            obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_pickle_shim)
        f = _lift(obj)
    except Exception as e:
        raise RuntimeError(f"Failed to load {path}. Ensure graph was created with torch.save(Data(...)). "
                           f"Original error: {e}") from e
    if "y" not in f or "timestep" not in f:
        raise RuntimeError(f"Failed to load {path}. Ensure graph was created with torch.save(Data(...)). "
                           "Original error: the object lacks y / timestep")
    return EllipticGraph(**f)


def to_device_graph(data, symmetrize_edges: bool = False, self_loops: bool = False, device="cuda"):
    """Move a loaded graph to the GPU and build its sorted views once: returns (data_on_device, edge_index, Graph) where
    `edge_index` is what the model must be called with (symmetrised like src/train_gnn.py:319-326 when asked) and
    the Graph is already registered in the cache under that tensor."""
    from .graph import build_graph, register_graph, symmetrize
    kw = {k: getattr(data, k).to(device) for k in _FIELDS if getattr(data, k, None) is not None}
    d = EllipticGraph(**kw)
    ei = symmetrize(d.edge_index) if symmetrize_edges else d.edge_index
    g = build_graph(ei, d.num_nodes, self_loops=self_loops)
    register_graph(ei, d.num_nodes, g, self_loops=self_loops)
    return d, ei, g
