"""Node-partitioned forward of the encode-process-decode model (BASELINE config 5).

Every rank holds one LocalGraph (owned nodes + ghost senders, the edges whose receiver it owns) and a
full copy of the weights.  A message-passing step runs the ordinary fused kernels on the local
graph; then the owners' fresh node rows replace the ghost copies (one all-to-all-v of bf16 rows, NCCL
over NVLink).  Because an edge lives with its receiver, the segment sum needs no communication and the
edge latent never moves.  The result on the owned nodes equals the unpartitioned forward (up to fp32
summation order at tile boundaries) -- unlike the reference's Cluster-GCN partitioning, which drops
the cut edges (graphphysics/dataset/dataset.py:258-264).

Training: `forward(..., save=True)` keeps the per-layer context, `backward` runs the ordinary backward
kernels on the local graph with the reverse exchange before every block (the gradient rows of the ghost
copies are added into their owners' rows, then zeroed) and sums the weight gradients over ranks: the
result equals the unpartitioned gradient of the same loss."""
from __future__ import annotations

import numpy as np
import torch

from ..graph import GraphCSR
from .halo import HaloPlan
from .partition import LocalGraph


class PartitionedEPD:
    def __init__(self, model, lg: LocalGraph, world: int, group=None):
        self.model, self.lg, self.group = model, lg, group
        self.engine = model.engine
        dev = self.engine.device
        self.plan = HaloPlan(lg, world, dev)
        self.csr = GraphCSR(torch.from_numpy(lg.edge_index_local).to(dev), lg.num_local)
        self.local_ids = torch.from_numpy(np.concatenate([lg.owned, lg.ghosts])).to(dev)
        self.edge_ids = torch.from_numpy(lg.edge_ids).to(dev)

    @torch.no_grad()
    def forward(self, x_global: torch.Tensor, edge_attr_global: torch.Tensor, save: bool = False):
        """Inputs are the (normalised) global node / edge features; every rank reads only its own rows.
        Returns the model output on this rank's owned nodes, in ascending global id (and, with
        save=True, the context `backward` needs)."""
        x_loc = x_global[self.local_ids].contiguous()
        ea_loc = edge_attr_global[self.edge_ids].contiguous()
        out, _, ctx = self.engine.forward(x_loc, ea_loc, self.csr, save=save,
                                          after_block=lambda x: self.plan.exchange_(x, self.group))
        out = out[: self.lg.num_owned]
        return (out, ctx) if save else out

    @torch.no_grad()
    def backward(self, ctx, d_out_owned: torch.Tensor) -> torch.Tensor:
        """d_out_owned = d loss / d output on the owned nodes (the loss is a sum over all ranks' owned
        nodes).  Leaves the gradient of the whole loss w.r.t. every parameter in engine.gflat on every
        rank and returns it."""
        import torch.distributed as dist
        d_out = torch.zeros((self.lg.num_local, d_out_owned.shape[1]), dtype=torch.float32, device=d_out_owned.device)
        d_out[: self.lg.num_owned] = d_out_owned
        self.engine.backward(ctx, d_out, before_block=lambda dx: self.plan.exchange_grad_(dx, self.group))
        dist.all_reduce(self.engine.gflat, op=dist.ReduceOp.SUM, group=self.group)
        return self.engine.gflat
