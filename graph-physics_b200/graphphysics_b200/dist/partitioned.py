"""Node-partitioned forward of the encode-process-decode model (BASELINE config 5).

Every rank holds one LocalGraph (owned nodes + ghost senders, the edges whose receiver it owns) and a
full copy of the weights.  A message-passing step runs the ordinary fused kernels on the local
graph; then the owners' fresh node rows replace the ghost copies (one all-to-all-v of bf16 rows, NCCL
over NVLink).  Because an edge lives with its receiver, the segment sum needs no communication and the
edge latent never moves.  The result on the owned nodes equals the unpartitioned forward (up to fp32
summation order at tile boundaries) -- unlike the reference's Cluster-GCN partitioning, which drops
the cut edges (graphphysics/dataset/dataset.py:258-264).

Forward / roll-out only in this round; training across partitions needs the reverse exchange (ghost
gradients added into their owners) and is listed as next in DESIGN.md."""
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
    def forward(self, x_global: torch.Tensor, edge_attr_global: torch.Tensor) -> torch.Tensor:
        """Inputs are the (normalised) global node / edge features; every rank reads only its own rows.
        Returns the model output on this rank's owned nodes, in ascending global id."""
        x_loc = x_global[self.local_ids].contiguous()
        ea_loc = edge_attr_global[self.edge_ids].contiguous()
        out, _, _ = self.engine.forward(x_loc, ea_loc, self.csr, save=False,
                                        after_block=lambda x: self.plan.exchange_(x, self.group))
        return out[: self.lg.num_owned]
