"""B200-native GNN message-passing hot path (GCN / SAGE / SAGE-ResBN / GAT).

Drop-in for the conv layers and nets that `/root/reference/src/models/gnn.py` builds over the
Elliptic transaction graph: same constructors, `forward(x, edge_index, t_idx)` signature and
state-dict names, computed by hand-written sm_100a kernels behind a C-ABI shared library
(`include/egnn_b200.h`, `libegnn_b200.so`).  CUDA only; no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401
from .graph import (Graph, GraphCache, ablate_hubs, build_graph, cached_graph, drop_edges,  # noqa: F401
                    register_graph, symmetrize)
from .nn import GATConv, GCNConv, SAGEConv  # noqa: F401
from .models import GATNet, GCNNet, SAGENet, SAGEResBNNet, build_model  # noqa: F401
from . import ingest, metrics, ops, synthetic  # noqa: F401
from .loader import Batch, NeighborLoader  # noqa: F401
from .ops import append_scalar_time, make_loss_fn  # noqa: F401
