"""NumPy / pure-Python oracle for the ingestion join  --  TEST INFRASTRUCTURE ONLY.

Restates, for checking `egnn_txid_join` / `egnn_temporal_masks` bit-for-bit:
  * txId -> row index, unknown-endpoint and cross-timestep edge filter
    /root/reference/src/data/dataset_elliptic.py:190-245
  * make_temporal_masks            /root/reference/src/data/dataset_elliptic.py:268-290
PINNED: tests/golden/make_ingest_golden.py runs the reference's own `load_elliptic_as_graph` / `make_temporal_masks`
(with `torch_geometric.data.Data` replaced by an attribute bag -- the loader only uses it as a container) on seeded
CSV tables and commits its outputs; tests/test_oracle_ingest.py holds this restatement to them.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def join_edges(tx_ids: np.ndarray, timestep: np.ndarray, e_src_tx: np.ndarray, e_dst_tx: np.ndarray):
    """-> (edge_index int64 [2, E_kept], mapped, kept)."""
    tx_to_idx = {int(tx): i for i, tx in enumerate(tx_ids)}          # :195-196 (a repeated txId keeps its LAST row)
    src, dst = [], []
    for s, d in zip(e_src_tx.tolist(), e_dst_tx.tolist()):           # :221-232, CSV order
        if s in tx_to_idx and d in tx_to_idx:
            src.append(tx_to_idx[s])
            dst.append(tx_to_idx[d])
    src_idx, dst_idx = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
    mapped = len(src_idx)
    if mapped:                                                       # :235-241
        same_t = timestep[src_idx] == timestep[dst_idx]
        src_idx, dst_idx = src_idx[same_t], dst_idx[same_t]
    return np.stack([src_idx, dst_idx]).astype(np.int64).reshape(2, -1), mapped, len(src_idx)


def temporal_masks(y: np.ndarray, t: np.ndarray, t_train_end: int, t_val_end: int, train_window_k=None):
    labeled = y >= 0
    train = (t <= t_train_end) & labeled
    val = (t > t_train_end) & (t <= t_val_end) & labeled
    test = (t > t_val_end) & labeled
    if train_window_k is not None:
        t_lo = max(1, t_train_end - train_window_k + 1)
        train = (t >= t_lo) & (t <= t_train_end) & labeled
    return train, val, test
