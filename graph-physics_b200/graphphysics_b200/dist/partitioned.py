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


class _HaloExchange(torch.autograd.Function):
    """Ghost rows of the node latent <- their owners' rows (HaloPlan.exchange_); the backward is the transpose (the
    ghost copies' gradient rows are added into their owners' rows, the ghost rows zeroed)."""

    @staticmethod
    def forward(ctx, x, plan: HaloPlan, group):
        ctx.plan, ctx.group = plan, group
        x = x.clone()
        plan.exchange_(x, group)
        return x

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous().clone()
        ctx.plan.exchange_grad_(d, ctx.group)
        return d, None, None


class PartitionedETD:
    """Node-partitioned EncodeTransformDecode ("the Transformer uses the same halo").

    An attention row lives with its query node: `lg` is the LocalGraph of the FLIPPED edge list (build_local_graph on
    edge_index[[1, 0]]), so a rank holds every stored (row = owned query, col = key) entry of its rows and ghost copies
    of the remote keys.  Every block runs the ordinary kernels on the local rows (k / v of the ghosts are recomputed
    from their latent here; what a block computes FOR a ghost row is discarded); before each block but the first the
    owners' fresh latent rows replace the ghost copies (fp32 residual stream, one all-to-all-v).  The encoder is
    row-wise, so the ghosts' first latent needs no exchange.  Autograd runs the transpose exchanges; parameter
    gradients are summed over ranks in `reduce_gradients`.  The result on the owned nodes equals the unpartitioned model."""

    def __init__(self, model, lg: LocalGraph, world: int, group=None):
        from ..graph import get_csr
        self.model, self.lg, self.group = model, lg, group
        dev = next(model.parameters()).device
        self.plan = HaloPlan(lg, world, dev)
        local = np.ascontiguousarray(lg.edge_index_local[::-1])          # back to (row = query, col = key)
        self.edge_index_local = torch.from_numpy(local).to(dev)
        self.csr = get_csr(self.edge_index_local, lg.num_local)
        self.local_ids = torch.from_numpy(np.concatenate([lg.owned, lg.ghosts])).to(dev)

    def forward(self, x_global: torch.Tensor, pos_global: torch.Tensor = None) -> torch.Tensor:
        """Model output on this rank's owned nodes (ascending global id); differentiable."""
        from .. import dense, variants
        m = self.model
        terms = 3 if m.precision == "tight" else 1
        x = x_global[self.local_ids].float().contiguous()
        pos = pos_global[self.local_ids].contiguous() if (pos_global is not None and m.use_rope_embeddings) else None
        if m.use_rope_embeddings and pos is None:
            raise ValueError("use_rope_embeddings=True requires 'pos' attribute in the input graph.")
        if m.act != "relu":
            enc = lambda seq, t: variants.mlp_seq(seq, t, m.act, terms)
        else:
            enc = lambda seq, t: dense.mlp4(seq, t, terms=terms)
        if not m.only_processor:
            x = enc(m.nodes_encoder, x)
        prev_x = x
        for i, block in enumerate(m.processor_list):
            if i > 0:
                x = _HaloExchange.apply(x, self.plan, self.group)
            prev_x = x
            x = block(x, self.csr, pos=pos)
        if m.temporal_block is not None:                       # its values come from the last block's output: ghosts too
            x = m.temporal_block(prev_x, _HaloExchange.apply(x, self.plan, self.group), self.csr)
        x = x[: self.lg.num_owned]
        return x if m.only_processor else enc(m.decode_module, x)

    def reduce_gradients(self) -> None:
        """Sum the parameter gradients over the ranks (every rank then holds the gradient of the whole loss)."""
        import torch.distributed as dist
        for p in self.model.parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
