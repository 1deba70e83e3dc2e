"""Halo exchange of boundary node rows between ranks.

Per message-passing step: pack the rows peers need (`gather_rows`), one all-to-all-v over the
process group (NCCL over NVLink on GPUs, gloo in the CPU tests), write the received rows into the
ghost slots (`scatter_rows`).  Send and receive lists enumerate the same global ids in ascending
order on both sides, so no ids travel -- only rows."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist

from .partition import LocalGraph


class HaloPlan:
    def __init__(self, lg: LocalGraph, world: int, device):
        self.world = world
        self.send_counts = [len(lg.send.get(q, ())) for q in range(world)]
        self.recv_counts = [len(lg.recv.get(q, ())) for q in range(world)]
        cat = lambda d: torch.cat([torch.as_tensor(d[q], dtype=torch.long) for q in range(world) if q in d]) \
            if d else torch.zeros(0, dtype=torch.long)
        self.send_idx = cat(lg.send).to(device)
        self.recv_idx = cat(lg.recv).to(device)

    def exchange_(self, x: torch.Tensor, group=None) -> None:
        """x[recv_idx] <- rows x[send_idx] of the owning ranks (in place on the ghost rows of x)."""
        h = x.shape[1]
        out = x.index_select(0, self.send_idx).contiguous()
        inp = torch.empty((int(sum(self.recv_counts)), h), dtype=x.dtype, device=x.device)
        if x.dtype == torch.bfloat16 and dist.get_backend(group) == "gloo":
            o32, i32 = out.float(), inp.float()               # gloo has no bf16 all-to-all
            dist.all_to_all_single(i32, o32, self.recv_counts, self.send_counts, group=group)
            inp = i32.to(x.dtype)
        else:
            dist.all_to_all_single(inp, out, self.recv_counts, self.send_counts, group=group)
        x.index_copy_(0, self.recv_idx, inp)

    def exchange_grad_(self, dx: torch.Tensor, group=None) -> None:
        """Transpose of `exchange_`: the gradient rows of the ghost copies travel back to their owners
        and are added to the owners' rows; the ghost rows are then zeroed (what a rank computed for a
        ghost before the exchange overwrote it has no consumer).  fp32, in place."""
        h = dx.shape[1]
        out = dx.index_select(0, self.recv_idx).contiguous()
        inp = torch.empty((int(sum(self.send_counts)), h), dtype=dx.dtype, device=dx.device)
        dist.all_to_all_single(inp, out, self.send_counts, self.recv_counts, group=group)
        dx.index_add_(0, self.send_idx, inp)           # a row sent to several peers collects all of them
        dx.index_fill_(0, self.recv_idx, 0.0)

