"""Graph-structure oracle (numpy, int64, exact) — TEST INFRASTRUCTURE ONLY.

Restates the DGL 0.5 graph operations the reference's preprocessing relies on
(DGL itself is not available offline):

* ``to_bidirected``      src/no-sampling/run.py:137   (add reverse edges, then to_simple)
* ``remove_self_loop``   src/no-sampling/run.py:143
* ``add_self_loop``      src/no-sampling/run.py:143   (appends (i,i), i=0..N-1)
* ``create_formats_``    src/no-sampling/run.py:146, src/ogbn-proteins/gat.py:66
                         (CSR by src + CSC by dst, stable counting sort => inside a
                         row edges appear in increasing edge id)
* degree tables          src/ogbn-proteins/gat.py:64  (out_degrees().float().clamp(min=1))

plus the partition maps of the multi-GPU path (no reference implementation exists;
the contract is the one written in DESIGN.md section "Partitioning").
"""
from __future__ import annotations

import numpy as np


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def to_bidirected(src, dst, n_nodes):
    """Add reverse edges, drop duplicates; result sorted by (src, dst)."""
    src, dst = _i64(src), _i64(dst)
    s = np.concatenate([src, dst])
    d = np.concatenate([dst, src])
    key = np.unique(s * np.int64(n_nodes) + d)  # sorted ascending == (src, dst) order
    return key // n_nodes, key % n_nodes


def remove_self_loop(src, dst):
    src, dst = _i64(src), _i64(dst)
    keep = src != dst
    return src[keep], dst[keep]


def add_self_loop(src, dst, n_nodes):
    src, dst = _i64(src), _i64(dst)
    loop = np.arange(n_nodes, dtype=np.int64)
    return np.concatenate([src, loop]), np.concatenate([dst, loop])


def build_csr(key, other, n_rows, sort_neighbours=False):
    """Stable COO -> CSR keyed on ``key``: inside a row, increasing edge id (DGL's counting sort) or, with
    ``sort_neighbours``, increasing (neighbour id, edge id) — the canonical order of ``bot_b200.Graph``.

    Returns (indptr[n_rows+1], indices[E] = other in row order, eid[E]).
    """
    key, other = _i64(key), _i64(other)
    if sort_neighbours:
        eid = np.lexsort((np.arange(key.shape[0]), other, key)).astype(np.int64)
    else:
        eid = np.argsort(key, kind="stable").astype(np.int64)
    counts = np.bincount(key, minlength=n_rows).astype(np.int64)
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, other[eid], eid


def canonical_edge_ids(src, dst):
    """Canonical position of every edge id: rank of (dst, src, edge id) — ``bot_b200.Graph.canonical_edge_ids``."""
    src, dst = _i64(src), _i64(dst)
    perm = np.lexsort((np.arange(src.shape[0]), src, dst))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.shape[0])
    return inv.astype(np.int64)


def build_formats(src, dst, n_src, n_dst, canonical=False):
    """All structure arrays the device graph must reproduce bit-exactly: ``botgat_graph_create`` on the COO as
    given (DGL's order), or — ``canonical`` — ``bot_b200.Graph``'s default, neighbour lists sorted by id."""
    src, dst = _i64(src), _i64(dst)
    in_indptr, in_indices, in_eid = build_csr(dst, src, n_dst, canonical)     # CSC: rows = dst
    out_indptr, out_indices, out_eid = build_csr(src, dst, n_src, canonical)  # CSR: rows = src
    return {
        "in_indptr": in_indptr, "in_indices": in_indices, "in_eid": in_eid,
        "out_indptr": out_indptr, "out_indices": out_indices, "out_eid": out_eid,
        "in_deg": np.diff(in_indptr), "out_deg": np.diff(out_indptr),
    }


def deg_scale(deg, power):
    """clamp(deg,1)^power in fp32, the way torch does it on a float tensor
    (src/no-sampling/models.py:501-502, 551-552)."""
    d = np.maximum(np.asarray(deg, dtype=np.float32), np.float32(1.0))
    if power == -0.5:
        return (np.float32(1.0) / np.sqrt(d)).astype(np.float32)
    if power == 0.5:
        return np.sqrt(d).astype(np.float32)
    return np.power(d, np.float32(power)).astype(np.float32)


# --------------------------------------------------------------------------
# 1-D destination-row partition (multi-GPU path)
# --------------------------------------------------------------------------

def partition_bounds(in_indptr, n_parts):
    """Contiguous dst-row ranges balanced on edge count.

    bounds[p] = first row r with in_indptr[r] >= floor(p*E/P); bounds[0]=0,
    bounds[P]=N.  Rank p owns rows [bounds[p], bounds[p+1]).
    """
    in_indptr = _i64(in_indptr)
    n = in_indptr.shape[0] - 1
    e = int(in_indptr[-1])
    bounds = np.zeros(n_parts + 1, dtype=np.int64)
    for p in range(1, n_parts):
        target = (p * e) // n_parts
        bounds[p] = np.searchsorted(in_indptr, target, side="left")
    bounds[n_parts] = n
    bounds[1:n_parts] = np.minimum(bounds[1:n_parts], n)
    return np.maximum.accumulate(bounds)


def partition_local(src, dst, n_nodes, bounds, rank):
    """Local edge set and maps of ``rank`` for a homogeneous graph.

    Local source numbering = [owned rows (global id - lo) | halo rows sorted by
    global id].  Local edges keep their relative (edge-id) order.
    Returns dict with: lo, hi, edge_gid (global eid of each local edge),
    lsrc, ldst (local ids), halo_gid, halo_owner, recv_counts[P], send_lists
    (for every peer q: the global ids *this rank* must send to q, sorted).
    """
    src, dst = _i64(src), _i64(dst)
    bounds = _i64(bounds)
    n_parts = bounds.shape[0] - 1
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    mine = np.nonzero((dst >= lo) & (dst < hi))[0].astype(np.int64)
    s, d = src[mine], dst[mine]
    owned = (s >= lo) & (s < hi)
    halo_gid = np.unique(s[~owned])
    lsrc = np.where(owned, s - lo, (hi - lo) + np.searchsorted(halo_gid, s))
    ldst = d - lo
    halo_owner = np.searchsorted(bounds, halo_gid, side="right") - 1
    recv_counts = np.bincount(halo_owner, minlength=n_parts).astype(np.int64)
    # what this rank must send: rows it owns that some peer q references
    send_lists = []
    for q in range(n_parts):
        if q == rank:
            send_lists.append(np.zeros(0, dtype=np.int64))
            continue
        qlo, qhi = int(bounds[q]), int(bounds[q + 1])
        sel = (dst >= qlo) & (dst < qhi) & (src >= lo) & (src < hi)
        send_lists.append(np.unique(src[sel]))
    return {
        "lo": lo, "hi": hi, "edge_gid": mine, "lsrc": lsrc.astype(np.int64),
        "ldst": ldst.astype(np.int64), "halo_gid": halo_gid,
        "halo_owner": halo_owner.astype(np.int64), "recv_counts": recv_counts,
        "send_lists": send_lists,
    }


# --------------------------------------------------------------------------
# synthetic graphs of the BASELINE.json shapes
# --------------------------------------------------------------------------

def synthetic_coo(n_nodes, n_edges, seed=0, power_law=0.0):
    """Random COO.  ``power_law`` > 0 draws dst proportional to rank^-power_law."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n_nodes, size=n_edges, dtype=np.int64)
    if power_law > 0:
        w = np.arange(1, n_nodes + 1, dtype=np.float64) ** (-power_law)
        cdf = np.cumsum(w / w.sum())
        dst = np.minimum(np.searchsorted(cdf, rng.random(n_edges)), n_nodes - 1).astype(np.int64)
    else:
        dst = rng.integers(0, n_nodes, size=n_edges, dtype=np.int64)
    return src, dst
