"""Ingestion oracle (oracle/ingest_np.py) held to the outputs of the REFERENCE's own loader
(tests/golden/ingest_golden.pt, written by tests/golden/make_ingest_golden.py from
/root/reference/src/data/dataset_elliptic.py), plus the PyG-free graph.pt reader on the host."""
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ingest_np as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_golden.pt")


@pytest.fixture(scope="module")
def cases():
    return torch.load(GOLD, weights_only=False)


def test_join_matches_reference_loader(cases):
    assert len(cases) >= 5
    for c in cases:
        # node order and txIds as the loader saw them (row i of the features CSV)
        ei, mapped, kept = O.join_edges(c["tx_ids"].numpy(), c["timestep"].numpy(), c["e_src_tx"].numpy(),
                                        c["e_dst_tx"].numpy())
        assert ei.dtype == np.int64 and ei.shape == tuple(c["edge_index"].shape)
        assert np.array_equal(ei, c["edge_index"].numpy()), c["spec"]
        assert kept == c["meta"]["num_edges"]
        assert f"mapped={mapped} " in c["log"] and f"same_t={kept} " in c["log"]


def test_masks_match_reference(cases):
    for c in cases:
        for name, m in c["masks"].items():
            a, b, k = m["args"]
            tr, va, te = O.temporal_masks(c["y"].numpy(), c["timestep"].numpy(), a, b, k)
            assert np.array_equal(tr, m["train"].numpy()) and np.array_equal(va, m["val"].numpy())
            assert np.array_equal(te, m["test"].numpy()), (c["spec"], name)


def test_reference_own_mask_fixture():
    """/root/reference/tests/test_masks_and_metrics.py:8-18 -- the 5-node path graph."""
    y = np.array([0, 1, -1, 0, 1])
    t = np.array([1, 1, 2, 3, 3])
    tr, va, te = O.temporal_masks(y, t, 1, 2)
    assert tr.tolist() == [True, True, False, False, False]
    assert va.tolist() == [False] * 5
    assert te.tolist() == [False, False, False, True, True]


def test_duplicate_txid_keeps_last_row():
    tx = np.array([5, 9, 5], dtype=np.int64)
    t = np.array([1, 1, 1], dtype=np.int64)
    ei, mapped, kept = O.join_edges(tx, t, np.array([5, 9], dtype=np.int64), np.array([9, 5], dtype=np.int64))
    assert ei.tolist() == [[2, 1], [1, 2]] and mapped == kept == 2


def test_load_cached_without_pyg(tmp_path):
    """A `graph.pt` whose pickle names torch_geometric classes loads without torch_geometric: the container classes
    become attribute bags and the tensors are lifted out (src/train_gnn.py:50-64 contract: x / edge_index / y /
    timestep / masks)."""
    from egnn_b200 import ingest
    # fabricate the PyG class paths for the WRITER only (Data -> _store: GlobalStorage -> _mapping: dict)
    mods = {}
    for name in ("torch_geometric", "torch_geometric.data", "torch_geometric.data.data", "torch_geometric.data.storage"):
        mods[name] = types.ModuleType(name)

    class GlobalStorage:
        pass

    class Data:
        pass

    GlobalStorage.__module__, GlobalStorage.__qualname__ = "torch_geometric.data.storage", "GlobalStorage"
    Data.__module__, Data.__qualname__ = "torch_geometric.data.data", "Data"
    mods["torch_geometric.data.storage"].GlobalStorage = GlobalStorage
    mods["torch_geometric.data.data"].Data = Data
    tensors = {"x": torch.randn(6, 3), "edge_index": torch.tensor([[0, 1, 2], [1, 2, 3]]), "y": torch.tensor([0, 1, -1, 0, 1, -1]),
               "timestep": torch.tensor([1, 1, 2, 2, 3, 3]), "train_mask": torch.tensor([1, 1, 0, 0, 0, 0]).bool(),
               "val_mask": torch.tensor([0, 0, 0, 1, 0, 0]).bool(), "test_mask": torch.tensor([0, 0, 0, 0, 1, 0]).bool()}
    st = GlobalStorage()
    st.__dict__["_mapping"] = dict(tensors)
    d = Data()
    d.__dict__["_store"] = st
    d.__dict__["_edge_attr_cls"] = None
    saved = dict(sys.modules)
    try:
        sys.modules.update(mods)
        torch.save(d, tmp_path / "graph.pt")
    finally:
        for k in mods:
            sys.modules.pop(k, None)
        sys.modules.update({k: v for k, v in saved.items() if k in mods})
    assert "torch_geometric" not in sys.modules
    g = ingest.load_cached(str(tmp_path))
    for k, v in tensors.items():
        assert torch.equal(getattr(g, k), v), k
    # a plain dict of tensors works too; a missing file raises the reference's message
    os.makedirs(tmp_path / "d2")
    torch.save(tensors, tmp_path / "d2" / "graph.pt")
    assert torch.equal(ingest.load_cached(str(tmp_path / "d2")).edge_index, tensors["edge_index"])
    with pytest.raises(RuntimeError, match="Failed to load"):
        ingest.load_cached(str(tmp_path / "nope"))
