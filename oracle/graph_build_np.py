"""NumPy oracle for the graph-build integer path  --  TEST INFRASTRUCTURE ONLY.

Restates, for checking the sm_100a sort/scan/histogram kernels bit-for-bit:
  * edge symmetrisation        /root/reference/src/train_gnn.py:319-326
  * PyG add_remaining_self_loops / gcn_norm (third-party
    torch_geometric 2.5.3, not vendored -> PARITY UNPINNED, see oracle/pyg_restated.py)
  * the stable destination-sorted (CSR) and source-sorted (CSC) views that reproduce
    the CPU `scatter_add_` / `index_select`-backward summation order
  * the Philox4x32-10 dropout keep-mask the CUDA path draws (so the oracle can be fed the
    very same mask)

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def symmetrize(ei: np.ndarray) -> np.ndarray:
    """cat([ei, ei.flip(0)], dim=1): second half = reversed pairs, same order, no dedup."""
    return np.concatenate([ei, ei[::-1]], axis=1)


def add_remaining_self_loops(ei: np.ndarray, n: int) -> np.ndarray:
    keep = ei[0] != ei[1]
    loops = np.arange(n, dtype=ei.dtype)
    return np.concatenate([ei[:, keep], np.stack([loops, loops])], axis=1)


def sorted_view(ei: np.ndarray, n: int, by: int):
    """Stable counting sort of the edge list by ei[by] -> (ptr[n+1] i32, other[E] i32, eid[E] i32).

    by=1: CSR by destination, `other` = source of each incoming edge, in original edge order.
    by=0: CSC by source, `other` = destination of each outgoing edge, in original edge order.
    """
    key = ei[by]
    perm = np.argsort(key, kind="stable")
    counts = np.bincount(key, minlength=n)
    ptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(counts, out=ptr[1:])
    return ptr, ei[1 - by][perm].astype(np.int32), perm.astype(np.int32)


def in_degree(ei: np.ndarray, n: int) -> np.ndarray:
    return np.bincount(ei[1], minlength=n).astype(np.int32)


def gcn_norm(ei: np.ndarray, n: int):
    """-> (ei2 [2,E2] with self loops, w [E2] float32) bit-equal to torch CPU:
    dis = rn(1/rn(sqrt(deg))), w = rn(dis[src]*dis[dst])."""
    ei2 = add_remaining_self_loops(ei, n)
    deg = np.bincount(ei2[1], minlength=n).astype(np.float32)
    with np.errstate(divide="ignore"):
        dis = (np.float32(1.0) / np.sqrt(deg)).astype(np.float32)
    dis[np.isinf(dis)] = 0
    w = (dis[ei2[0]] * np.float32(1.0)) * dis[ei2[1]]
    return ei2, w.astype(np.float32), dis


# ----------------------------------------------------------------------------
# Philox4x32-10 keep-mask (must match csrc/philox.cuh bit for bit)
# ----------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (counters) with scalar keys."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    mask32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = _M0 * c0.astype(np.uint64)
        p1 = _M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask32).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask32).astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        with np.errstate(over="ignore"):
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dropout_threshold(p: float) -> int:
    """keep iff u16 lane >= thr;  thr = min(floor(p * 2^16), 2^16-1)."""
    return int(min(int(np.floor(float(p) * 65536.0)), 65535))


def dropout_keep_mask(seed: int, layer: int, row0: int, n: int, f: int, p: float) -> np.ndarray:
    """uint8 [n, f]; element (r, c) uses Philox counter (row0+r lo, row0+r hi, c//8, layer), key (seed lo, seed hi)
    and the 16-bit lane c%8 of the four output words (even lane = low half, odd lane = high half of word (c%8)//2)."""
    rows = np.arange(row0, row0 + n, dtype=np.uint64)
    nb = (f + 7) // 8
    r_lo = np.repeat((rows & np.uint64(0xFFFFFFFF)).astype(np.uint32), nb)
    r_hi = np.repeat((rows >> np.uint64(32)).astype(np.uint32), nb)
    cb = np.tile(np.arange(nb, dtype=np.uint32), n)
    lay = np.full(n * nb, layer, dtype=np.uint32)
    o = philox4x32_10(r_lo, r_hi, cb, lay, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lanes = []
    for w in o:
        lanes += [w & np.uint32(0xFFFF), w >> np.uint32(16)]
    u = np.stack(lanes, axis=1).reshape(n, nb * 8)[:, :f]
    return (u >= np.uint32(dropout_threshold(p))).astype(np.uint8)
