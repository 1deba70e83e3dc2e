"""Import shim for the UNMODIFIED reference modules (TEST INFRASTRUCTURE, build container only).

`/root/reference` needs torch_geometric and dgl, which are not installable here.  This module
registers ~60 lines of stand-ins in sys.modules for the three third-party behaviours the hot path
touches, then the reference's own
graphphysics/models/{layers,processors,simulator}.py and utils/{loss,scheduler,nodetype}.py import
and run unchanged.  It is used by oracle/make_golden.py to produce tests/golden/*; nothing at test
or bench run time imports it (the GPU box has no /root/reference).

Restated third-party semantics:
  * torch_geometric.nn.MessagePassing.propagate(aggr="add", flow="source_to_target")
      (torch-geometric==2.6.1, requirements.txt:7): out = zeros(N,H).index_add_(0, edge_index[1], message(...));
      then update(out, x=..., phi=...)                              -- used at layers.py:926, 1031-1037
  * torch_geometric.data.Data: attribute bag, missing attribute -> None   -- simulator.py:169-174
  * dgl.sparse.spmatrix / bsddmm / SparseMatrix.softmax / bspmm (dgl, unpinned, README.md:122-129)
      -- used at layers.py:512-517, 550-554; processors.py:366
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn

import os

# The reference tree in the build container; on the GPU box (no /root/reference) the copy staged by
# oracle/build_ref.py into the git-ignored oracle/_ref/.
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = "/root/reference" if os.path.isdir("/root/reference/graphphysics") else _STAGED


def install(root: str = None):
    root = root or REFERENCE_ROOT
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_gp_shim", False):
        if root not in sys.path:
            sys.path.insert(0, root)
        return
    tg = types.ModuleType("torch_geometric")
    tg._gp_shim = True
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_data = types.ModuleType("torch_geometric.data")

    class MessagePassing(nn.Module):
        def __init__(self, aggr="add", flow="source_to_target", **kw):
            super().__init__()
            assert aggr == "add" and flow == "source_to_target"

        def propagate(self, edge_index, size=None, **kw):
            msg = self.message(edge_attr=kw["edge_attr"])
            n = size[1] if size is not None else kw["x"].size(0)
            out = msg.new_zeros((n, msg.size(1))).index_add_(0, edge_index[1], msg)
            return self.update(out, x=kw["x"], phi=kw.get("phi"))

    class TransformerConv(nn.Module):      # only constructed when DGL is missing; inert here
        def __init__(self, *a, **k):
            super().__init__()

    class Data:
        def __init__(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return None

    tg_nn.MessagePassing, tg_nn.TransformerConv = MessagePassing, TransformerConv
    tg_data.Data, tg_data.Batch = Data, Data
    tg.nn, tg.data = tg_nn, tg_data

    dgl = types.ModuleType("dgl")
    dglsp = types.ModuleType("dgl.sparse")

    class SparseMatrix:
        def __init__(self, row, col, val, shape):
            self.row, self.col, self.val, self.shape = row, col, val, shape

        def astype(self, dtype):
            return SparseMatrix(self.row, self.col, self.val.to(dtype), self.shape)

        def softmax(self):
            v = self.val if self.val.dim() > 1 else self.val[:, None]
            n = self.shape[0]
            mx = torch.full((n, v.shape[1]), -float("inf"), dtype=v.dtype, device=v.device)
            mx = mx.scatter_reduce(0, self.row[:, None].expand_as(v), v, reduce="amax", include_self=True)
            p = torch.exp(v - mx[self.row])
            den = torch.zeros((n, v.shape[1]), dtype=v.dtype, device=v.device).index_add_(0, self.row, p)
            out = p / den[self.row]
            return SparseMatrix(self.row, self.col, out if self.val.dim() > 1 else out[:, 0], self.shape)

    def spmatrix(indices, val=None, shape=None):
        row, col = indices[0], indices[1]
        if val is None:
            val = torch.ones(row.shape[0], device=row.device)
        return SparseMatrix(row, col, val, shape)

    def from_coo(row, col, val=None, shape=None):
        return spmatrix(torch.stack([row, col]), val, shape)

    def bsddmm(A, X1, X2):
        # val[e,b] = A.val[e] * sum_k X1[row_e, k, b] * X2[k, col_e, b]
        v = (X1[A.row] * X2.permute(1, 0, 2)[A.col]).sum(dim=1)
        return SparseMatrix(A.row, A.col, v * (A.val[:, None] if A.val.dim() == 1 else A.val), A.shape)

    def bspmm(A, X):
        # out[i,k,b] = sum_{e: row_e = i} A.val[e,b] * X[col_e, k, b]
        out = torch.zeros((A.shape[0],) + tuple(X.shape[1:]), dtype=X.dtype, device=X.device)
        return out.index_add_(0, A.row, A.val[:, None, :] * X[A.col])

    dglsp.SparseMatrix, dglsp.spmatrix, dglsp.from_coo = SparseMatrix, spmatrix, from_coo
    dglsp.bsddmm, dglsp.bspmm = bsddmm, bspmm
    dgl.sparse = dglsp

    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.data": tg_data,
                        "dgl": dgl, "dgl.sparse": dglsp})
    if root not in sys.path:
        sys.path.insert(0, root)


def import_reference(root: str = None):
    """Returns the reference modules (layers, processors, simulator, loss, scheduler, nodetype)."""
    install(root)
    import importlib
    names = ["graphphysics.models.layers", "graphphysics.models.processors", "graphphysics.models.simulator",
             "graphphysics.utils.loss", "graphphysics.utils.scheduler", "graphphysics.utils.nodetype"]
    return {n.split(".")[-1]: importlib.import_module(n) for n in names}
