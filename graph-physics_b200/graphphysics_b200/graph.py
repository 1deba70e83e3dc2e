"""Graph containers for the kernels: a PyG-like attribute bag and the receiver-sorted CSR layout.

The reference hands edge lists sorted by sender (PyG coalesce) and aggregates at the receiver
(`edge_index[1]`, graphphysics/models/layers.py:926, 1031-1037).  The kernels want every
receiver's edges contiguous, so a topology is converted once into `GraphCSR` and cached.
"""
from __future__ import annotations

from typing import Optional

import torch


class Data:
    """Minimal stand-in for torch_geometric.data.Data: an attribute bag where a missing attribute
    reads as None (the behaviour graphphysics/models/simulator.py:169-174 and
    processors.py:187 rely on).  A real PyG Data works everywhere this does."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return None

    def clone(self) -> "Data":
        return Data(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})

    def to(self, device, non_blocking: bool = False) -> "Data":
        return Data(**{k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
                       for k, v in self.__dict__.items()})

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])


Batch = Data


def collate(graphs) -> Data:
    """The union graph of a list of graphs, as PyG's DataLoader collate builds it: tensors whose first
    dimension is the node or edge count are concatenated on dim 0, `edge_index` / `face` / `tetra` on dim 1 with the node
    offset of their graph added, plus `batch` (graph id per node) and `ptr` (node offsets).  Other tensor attributes are
    stacked when every graph has them with the same shape, non-tensor attributes are gathered into a list.  The model code
    treats the result as one graph (a block-diagonal adjacency)."""
    graphs = list(graphs)
    if not graphs:
        raise ValueError("collate needs at least one graph")
    keys = [k for k in graphs[0].__dict__.keys() if all(getattr(g, k) is not None for g in graphs)]
    sizes = [int(g.x.shape[0]) if g.x is not None else int(g.pos.shape[0]) for g in graphs]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    out = {}
    for k in keys:
        vals = [getattr(g, k) for g in graphs]
        if not all(torch.is_tensor(v) for v in vals):
            out[k] = vals
        elif k in ("edge_index", "face", "tetra"):
            out[k] = torch.cat([v + o for v, o in zip(vals, offs)], dim=1)
        elif all(v.dim() >= 1 for v in vals) and all(v.shape[1:] == vals[0].shape[1:] for v in vals) and (
                all(v.shape[0] == n for v, n in zip(vals, sizes)) or
                ("edge_index" in keys and all(v.shape[0] == g.edge_index.shape[1] for v, g in zip(vals, graphs)))):
            out[k] = torch.cat(vals, dim=0)
        elif all(v.shape == vals[0].shape for v in vals):
            out[k] = torch.stack(vals, dim=0)
        else:
            out[k] = vals
    dev = graphs[0].x.device if graphs[0].x is not None else graphs[0].pos.device
    out["batch"] = torch.cat([torch.full((n,), i, dtype=torch.long, device=dev) for i, n in enumerate(sizes)])
    out["ptr"] = torch.tensor(offs, dtype=torch.long, device=dev)
    return Data(**out)


class GraphCSR:
    """Receiver-sorted edge layout of one topology (all int32, on the device of edge_index).

    perm_dst      : sorted position -> original edge id (stable sort by receiver)
    src, dst      : endpoints in sorted order
    rowptr_dst    : [N+1] receiver segments of the sorted list
    perm_src      : positions of the sorted list, stable-sorted by sender
    rowptr_src    : [N+1] sender segments of perm_src
    """

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, _persistent: bool = False):
        assert edge_index.dim() == 2 and edge_index.shape[0] == 2
        self.num_nodes = int(num_nodes)
        self.num_edges = int(edge_index.shape[1])
        self._inv_perm: Optional[torch.Tensor] = None
        if edge_index.is_cuda:
            # the native counting sorts of libgp_b200.so (gp_csr_from_coo): no ATen sort, no host round trip
            self._alloc_native(edge_index.device, _persistent)
            self.rebuild_(edge_index)
            return
        # host tensors (CPU tests, the partitioner): the same layout spelled with torch ops
        src, dst = edge_index[0].int(), edge_index[1].int()
        perm = torch.sort(dst, stable=True).indices
        src_s, dst_s = src[perm], dst[perm]
        self.perm_dst64 = perm
        self.perm_dst = perm.int()
        self.src, self.dst = src_s, dst_s
        self.rowptr_dst = self._rowptr(dst_s, num_nodes)
        src_sorted, perm_src = torch.sort(src_s, stable=True)
        self.perm_src = perm_src.int()
        self.rowptr_src = self._rowptr(src_sorted, num_nodes)
        # attention view: rows = senders (edge_index[0]); the row-sorted entry p is entry perm_src[p] of
        # the receiver-sorted list, so its column is dst[perm_src[p]]
        self.att_col = self.dst[self.perm_src.long()].contiguous()

    def _alloc_native(self, device, persistent: bool) -> None:
        from . import ops
        E, N = self.num_edges, self.num_nodes
        i32 = lambda n: torch.empty(n, dtype=torch.int32, device=device)
        self.perm_dst, self.src, self.dst, self.perm_src, self.att_col = i32(E), i32(E), i32(E), i32(E), i32(E)
        self.rowptr_dst, self.rowptr_src = i32(N + 1), i32(N + 1)
        self._ws = torch.empty(ops.csr_workspace_bytes(E, N), dtype=torch.uint8, device=device)
        # persistent layouts (replayed CUDA graphs) keep the previous edge_index and skip the rebuild when it repeats
        self._prev = torch.empty((2, E), dtype=torch.int64, device=device) if persistent else None
        self._state = torch.zeros(2, dtype=torch.int32, device=device) if persistent else None
        self._perm64 = torch.empty(E, dtype=torch.int64, device=device)

    def rebuild_(self, edge_index: torch.Tensor) -> "GraphCSR":
        """(Re)compute the layout in place from a CUDA edge_index of the same shape; every buffer keeps its address."""
        from . import ops
        assert edge_index.is_cuda and edge_index.shape == (2, self.num_edges)
        ei = edge_index if (edge_index.dtype == torch.int64 and edge_index.is_contiguous()) else edge_index.long().contiguous()
        out = dict(perm_dst=self.perm_dst, src=self.src, dst=self.dst, rowptr_dst=self.rowptr_dst, perm_src=self.perm_src,
                   rowptr_src=self.rowptr_src, att_col=self.att_col)
        ops.csr_from_coo(ei, self.num_nodes, out, self._ws, self._prev, self._state)
        self._perm64.copy_(self.perm_dst)          # int64 copy for torch indexing of the raw edge features
        self._inv_perm = None
        return self

    @property
    def perm_dst64(self) -> torch.Tensor:
        return self._perm64

    @perm_dst64.setter
    def perm_dst64(self, v: torch.Tensor) -> None:
        self._perm64 = v

    @staticmethod
    def _rowptr(sorted_ids: torch.Tensor, n: int) -> torch.Tensor:
        """rowptr[k] = number of ids < k, for ids already sorted.  searchsorted needs no host
        round trip (bincount does), so the layout can be rebuilt inside a CUDA graph."""
        bounds = torch.arange(n + 1, device=sorted_ids.device, dtype=sorted_ids.dtype)
        return torch.searchsorted(sorted_ids, bounds, right=False).int()

    @property
    def inv_perm_dst64(self) -> torch.Tensor:
        """original edge id -> sorted position."""
        if self._inv_perm is None:
            inv = torch.empty_like(self.perm_dst64)
            inv[self.perm_dst64] = torch.arange(self.num_edges, device=inv.device)
            self._inv_perm = inv
        return self._inv_perm


_CACHE: dict = {}
_PERSISTENT: dict = {}       # (E, N, device) -> GraphCSR with fixed buffers, rebuilt in place (CUDA-graph replay)


_CACHE_ENABLED = [True]


class no_csr_cache:
    """Inside this context every get_csr call re-sorts (used while a CUDA graph is being captured, so
    the replay recomputes the layout from the current contents of the static input buffers)."""

    def __enter__(self):
        self.old = _CACHE_ENABLED[0]
        _CACHE_ENABLED[0] = False

    def __exit__(self, *a):
        _CACHE_ENABLED[0] = self.old


def get_csr(edge_index: torch.Tensor, num_nodes: int) -> GraphCSR:
    """GraphCSR of `edge_index`, cached on (storage pointer, version, shape) so a static topology
    (roll-outs, repeated batches) is sorted once."""
    if not _CACHE_ENABLED[0]:
        if not edge_index.is_cuda:
            return GraphCSR(edge_index, num_nodes)
        # replayed steps: one layout object per shape, rebuilt in place from the current contents of the static input
        # buffer -- and not rebuilt at all (device-side comparison) while the topology repeats
        key = (int(edge_index.shape[1]), int(num_nodes), str(edge_index.device))
        g = _PERSISTENT.get(key)
        if g is None:
            g = _PERSISTENT[key] = GraphCSR(edge_index, num_nodes, _persistent=True)
            return g
        return g.rebuild_(edge_index)
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), int(num_nodes), str(edge_index.device))
    g = _CACHE.get(key)
    if g is None:
        if len(_CACHE) > 64:
            _CACHE.clear()
        g = GraphCSR(edge_index, num_nodes)
        # keep the tensor alive so the pointer cannot be recycled for a different graph
        g._keepalive = edge_index
        _CACHE[key] = g
    return g
