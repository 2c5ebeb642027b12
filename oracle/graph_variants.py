"""CPU oracle for the re-entrant graph variants  --  TEST INFRASTRUCTURE ONLY.

Restates, in plain torch CPU ops,
  * `build_edge_index_ablated`  (`/root/reference/src/analysis/hub_ablation.py:56-71`; the same code inline at
    `/root/reference/src/train_gnn.py:526-540`): total degree = bincount(src) + bincount(dst), the `int(frac * N)`
    highest-degree nodes are hubs, every edge touching a hub is removed, the survivors keep their order;
  * `drop_edges`                (`/root/reference/src/analysis/robustness.py:65-82`): `round(drop_frac * E)` edges are
    dropped, the kept ones are `edge_index[:, perm[drop_count:]]` for a permutation `perm` of the edge columns.

PINNED: `tests/golden/make_variants_golden.py` imports the reference's own two functions (their modules import
torch_geometric only for the `Data` type, stubbed) and commits what they return on seeded graphs
(`tests/golden/variants_golden.pt`); `tests/test_oracle_graph_variants.py` holds this file to those vectors bit for bit.

The one rule the reference leaves to torch: which of several nodes TIED at the k-th largest degree `torch.topk`
returns (unspecified, implementation-dependent).  `stable=True` resolves ties towards the lower node id (what the
CUDA path does and documents); whenever the k-th and (k+1)-th largest degrees differ, both give the reference's set.
Only `tests/` may import this module.
"""
from typing import Optional, Tuple

import torch


def ablate_hubs(edge_index: torch.Tensor, num_nodes: int, frac: float, stable: bool = True):
    """-> (edge_index_ablated, num_hubs, hub mask bool[N], untied)."""
    num_hubs = int(float(frac) * float(num_nodes))                          # hub_ablation.py:59-60
    ei = edge_index.detach().cpu()
    deg = torch.bincount(ei[0], minlength=num_nodes) + torch.bincount(ei[1], minlength=num_nodes)   # :62-64
    hubs = torch.zeros(num_nodes, dtype=torch.bool)
    untied = True
    if num_hubs > 0:                                                        # :66-68
        idx = (torch.sort(deg, descending=True, stable=True).indices[:num_hubs] if stable
               else torch.topk(deg, num_hubs).indices)
        hubs[idx] = True
        srt = torch.sort(deg, descending=True).values
        untied = num_hubs >= num_nodes or bool(srt[num_hubs - 1] != srt[num_hubs])
    mask = ~(hubs[ei[0]] | hubs[ei[1]])                                     # :69
    return ei[:, mask], num_hubs, hubs, untied


def drop_count(num_edges: int, drop_frac: float) -> int:
    """robustness.py:66-79: argument check, Python `round` (half to even), clamp, the all-edges error."""
    drop_frac = float(drop_frac)
    if drop_frac < 0 or drop_frac > 1:
        raise ValueError("drop_frac must be within [0, 1]")
    if drop_frac <= 0:
        return 0
    n = min(int(round(drop_frac * float(num_edges))), num_edges)
    if n and n >= num_edges:
        raise RuntimeError("Dropping all edges would leave an empty graph.")
    return n


def drop_edges(edge_index: torch.Tensor, drop_frac: float, perm: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, int]:
    """robustness.py:65-82 with the permutation as an argument (`None`: drawn from torch's global CPU generator exactly
    where the reference draws it, so `torch.manual_seed(s)` before the call reproduces the reference's result)."""
    n = drop_count(edge_index.size(1), drop_frac)
    if n == 0:
        return edge_index, 0
    if perm is None:
        perm = torch.randperm(edge_index.size(1), device=edge_index.device)   # :80
    return edge_index[:, perm[n:]], n                                          # :81-82
