"""Node partition of one large mesh and the halo maps for message passing across ranks
(BASELINE config 5).

The reference's "partitioning" is Cluster-GCN: METIS parts are trained as independent sub-meshes and
the cut edges are DROPPED (graphphysics/dataset/dataset.py:244-302, utils/torch_graph.py:108-135), so
its partitioned result differs from its unpartitioned one.  Here the partitioned forward equals the
unpartitioned forward: an edge lives on the rank that owns its receiver (the segment sum stays local
and complete, `e` never moves), each rank keeps ghost copies of the remote senders it needs, and after
every message-passing step the owners send the fresh rows of those nodes (halo exchange).

METIS is not available; the partitioner is a deterministic recursive coordinate bisection (balanced
parts, ties broken by node id).  All maps are integer arrays, bit-exact against oracle/gp_oracle.py."""
from __future__ import annotations

from typing import Dict, List

import numpy as np


def partition_nodes(pos: np.ndarray, num_parts: int) -> np.ndarray:
    """owner[n] in [0, num_parts): recursive coordinate bisection along the widest axis."""
    n = pos.shape[0]
    owner = np.zeros(n, np.int32)

    def rec(ids: np.ndarray, lo: int, parts: int) -> None:
        if parts == 1:
            owner[ids] = lo
            return
        p = pos[ids]
        axis = int(np.argmax(p.max(0) - p.min(0)))
        order = np.lexsort((ids, p[:, axis]))
        half = (len(ids) * (parts // 2)) // parts
        rec(ids[order[:half]], lo, parts // 2)
        rec(ids[order[half:]], lo + parts // 2, parts - parts // 2)

    rec(np.arange(n, dtype=np.int64), 0, num_parts)
    return owner


class LocalGraph:
    """One rank's share: local node order = [owned ascending by global id | ghosts ascending]."""

    def __init__(self, rank: int, owned: np.ndarray, ghosts: np.ndarray, edge_ids: np.ndarray,
                 edge_index_local: np.ndarray, send: Dict[int, np.ndarray], recv: Dict[int, np.ndarray]):
        self.rank = rank
        self.owned, self.ghosts = owned, ghosts
        self.edge_ids = edge_ids                    # global ids of the edges kept here (receiver owned)
        self.edge_index_local = edge_index_local    # (2, E_local) in local numbering
        self.send, self.recv = send, recv           # peer -> local row indices (ascending global id)

    @property
    def num_owned(self) -> int:
        return len(self.owned)

    @property
    def num_local(self) -> int:
        return len(self.owned) + len(self.ghosts)


def build_local_graphs(edge_index: np.ndarray, owner: np.ndarray, num_parts: int) -> List[LocalGraph]:
    src, dst = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    n = owner.shape[0]
    parts, luts = [], []
    for p in range(num_parts):
        owned = np.nonzero(owner == p)[0].astype(np.int64)
        eids = np.nonzero(owner[dst] == p)[0].astype(np.int64)
        s = src[eids]
        ghosts = np.unique(s[owner[s] != p]).astype(np.int64)
        lut = np.full(n, -1, np.int64)
        lut[owned] = np.arange(len(owned))
        lut[ghosts] = len(owned) + np.arange(len(ghosts))
        recv = {}
        for q in range(num_parts):
            gq = ghosts[owner[ghosts] == q]
            if q != p and len(gq):
                recv[q] = lut[gq].astype(np.int32)
        parts.append(LocalGraph(p, owned, ghosts, eids, np.stack([lut[s], lut[dst[eids]]]).astype(np.int64), {}, recv))
        luts.append(lut)
    for p in range(num_parts):
        glob = np.concatenate([parts[p].owned, parts[p].ghosts])
        for q, idx in parts[p].recv.items():
            parts[q].send[p] = luts[q][glob[idx]].astype(np.int32)
    return parts


def build_local_graph(edge_index: np.ndarray, owner: np.ndarray, num_parts: int, rank: int) -> LocalGraph:
    """The LocalGraph of ONE rank (same result as build_local_graphs(...)[rank]) in O(E) work: what every rank of a
    large job computes for itself."""
    src, dst = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    n = owner.shape[0]
    p = rank
    o_src, o_dst = owner[src], owner[dst]
    owned = np.nonzero(owner == p)[0].astype(np.int64)
    eids = np.nonzero(o_dst == p)[0].astype(np.int64)
    s = src[eids]
    ghosts = np.unique(s[owner[s] != p]).astype(np.int64)
    lut = np.full(n, -1, np.int64)
    lut[owned] = np.arange(len(owned))
    lut[ghosts] = len(owned) + np.arange(len(ghosts))
    recv, send = {}, {}
    go = owner[ghosts]
    for q in range(num_parts):
        if q == p:
            continue
        gq = ghosts[go == q]
        if len(gq):
            recv[q] = lut[gq].astype(np.int32)
        # rows this rank owns that rank q holds as ghosts: senders (owned here) of edges whose receiver q owns
        mine = np.unique(src[(o_src == p) & (o_dst == q)])
        if len(mine):
            send[q] = lut[mine].astype(np.int32)
    return LocalGraph(p, owned, ghosts, eids, np.stack([lut[s], lut[dst[eids]]]).astype(np.int64), send, recv)
