"""Oracle for the neighbour sampler (TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).

Plain-Python restatement of ``torch_geometric.loader.NeighborLoader`` as the reference configures it
(biomedkg/data_module.py:71-79 ``num_neighbors=[-1]``; :81-99 ``[30]*3``): homogeneous graph, ``replace=False``,
``directed=True``, seeds first; per hop every node added in the previous hop draws ``min(in_degree, fanout)`` of its
in-edges; the sampled edges are the batch's only edges (row = local source, col = local target); newly reached sources
are appended in order of first appearance.

PARITY UNPINNED for the random stream: torch_geometric == 2.5.3 / pyg-lib (pyproject.toml:6) are not installable here
and their C++ RNG cannot be reproduced, so *which* neighbours are drawn is this repository's own counter-based stream
(the same formulas as csrc/sampler.cu, restated below); the structural rules above are PyG's published behaviour and are
what tests check against brute force.
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15


def hash_u32(seed: int, idx: int) -> int:
    """csrc/common.cuh hash_u32 on Python ints."""
    z = (idx * _GOLD + seed) & _M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z >> 32


def draw(seed: int, hop: int, node: int, lane: int, attempt: int, deg: int) -> int:
    idx = (node << 32) | (attempt << 8) | lane
    r = hash_u32((seed + _GOLD * (hop + 1)) & _M64, idx)
    return (r * deg) >> 32


def pick_positions(seed: int, hop: int, node: int, deg: int, fanout: int):
    """Positions (ascending) within the node's in-edge list (sorted by source, ties by edge id) that are sampled."""
    k = deg if (fanout < 0 or deg < fanout) else fanout
    if k == deg:
        return list(range(deg))
    exclude = deg <= 2 * k
    m = deg - k if exclude else k
    cand = [draw(seed, hop, node, lane, 0, deg) for lane in range(m)]
    attempt = [0] * m
    while True:
        dup = [any(cand[j] == cand[i] for j in range(i)) for i in range(m)]
        if not any(dup):
            break
        for i in range(m):
            if dup[i]:
                attempt[i] += 1
                cand[i] = draw(seed, hop, node, i, attempt[i], deg)
    chosen = set(cand)
    return sorted(set(range(deg)) - chosen) if exclude else sorted(chosen)


def sample(edge_index: np.ndarray, num_nodes: int, seeds, num_neighbors, seed: int):
    """-> (n_id, local edge_index [2,T], e_id) exactly as biomedkg_b200.loader.NeighborSampler.sample returns them."""
    src, dst = edge_index[0].astype(np.int64), edge_index[1].astype(np.int64)
    order = np.argsort(dst * num_nodes + src, kind="stable")        # in-edges of a node: by source, ties by edge id
    dsts = dst[order]
    rowptr = np.searchsorted(dsts, np.arange(num_nodes + 1))
    nodes = [int(s) for s in seeds]
    local = {v: i for i, v in enumerate(nodes)}
    assert len(local) == len(nodes), "seeds must be distinct"
    rows, cols, eids = [], [], []
    begin, end = 0, len(nodes)
    for hop, fanout in enumerate(num_neighbors):
        hop_src = []
        for i in range(begin, end):
            v = nodes[i]
            b, e = int(rowptr[v]), int(rowptr[v + 1])
            for p in pick_positions(seed, hop, v, e - b, fanout):
                eid = int(order[b + p])
                hop_src.append(int(src[eid]))
                cols.append(i)
                eids.append(eid)
        for s in hop_src:                                            # new nodes in order of first appearance
            if s not in local:
                local[s] = len(nodes)
                nodes.append(s)
        rows.extend(local[s] for s in hop_src)
        begin, end = end, len(nodes)
    return (np.asarray(nodes, dtype=np.int64), np.asarray([rows, cols], dtype=np.int64).reshape(2, -1),
            np.asarray(eids, dtype=np.int64))


def check_structure(edge_index: np.ndarray, num_nodes: int, seeds, num_neighbors, n_id, sub, e_id):
    """Brute-force check of a sampled batch against PyG NeighborLoader's published STRUCTURAL rules, independent of any random
    stream (it never calls ``draw`` / ``pick_positions``): returns a list of violated rules (empty = the batch is valid).

      1. the seeds come first in ``n_id``, which has no duplicates;
      2. every sampled edge IS an original edge: (n_id[row], n_id[col]) == edge_index[:, e_id];
      3. no replacement: e_id has no duplicates;
      4. fan-out cap: a node expanded in a hop with fan-out f contributes exactly min(in_degree, f) in-edges (all of them for
         f = -1); a node is expanded in the hop after the one that added it, nodes added by the last hop are not expanded;
      5. new nodes are appended in order of first appearance as a source within the hop that adds them;
      6. edges are grouped by hop and, inside a hop, by target in frontier order."""
    src, dst = edge_index[0].astype(np.int64), edge_index[1].astype(np.int64)
    n_id, sub, e_id = np.asarray(n_id), np.asarray(sub), np.asarray(e_id)
    bad = []
    S = len(seeds)
    if list(n_id[:S]) != [int(s) for s in seeds] or len(set(n_id.tolist())) != len(n_id):
        bad.append("1 seeds-first / distinct nodes")
    if sub.shape[1] != len(e_id) or (len(e_id) and (not np.array_equal(n_id[sub[0]], src[e_id]) or not np.array_equal(n_id[sub[1]], dst[e_id]))):
        bad.append("2 sampled edges are original edges")
    if len(set(e_id.tolist())) != len(e_id):
        bad.append("3 no replacement")
    indeg = np.bincount(dst, minlength=num_nodes)
    begin, end, pos, known = 0, S, 0, S
    for hop, f in enumerate(num_neighbors):
        hop_begin = pos
        for t in range(begin, end):                                  # frontier of this hop, in order
            want = int(indeg[n_id[t]]) if (f < 0 or indeg[n_id[t]] < f) else f
            got = 0
            while pos < sub.shape[1] and sub[1, pos] == t:
                got += 1
                pos += 1
            if got != want:
                bad.append(f"4/6 hop {hop} target {t}: {got} edges, expected {want}")
                return bad
        first_seen = []
        for r in sub[0, hop_begin:pos].tolist():
            if r >= known and r not in first_seen:
                first_seen.append(r)
        if first_seen != list(range(known, known + len(first_seen))):
            bad.append(f"5 first-appearance order in hop {hop}")
        begin, end = end, known + len(first_seen)
        known = end
    if pos != sub.shape[1] or known != len(n_id):
        bad.append("6 trailing edges / nodes not accounted for")
    return bad
