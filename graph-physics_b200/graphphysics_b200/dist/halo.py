"""Halo exchange of boundary node rows between ranks.

Per message-passing step: pack the rows peers need (gp_halo_pack), one all-to-all-v over the process group (NCCL over
NVLink on GPUs, gloo in the CPU tests), write the received rows into the ghost slots (gp_halo_unpack).  The backward
is the transpose: ghost gradient rows travel back and are ADDED into their owners' rows in a fixed order
(gp_halo_unpack_add: one thread owns a destination row chunk and walks its contributions ascending -- no atomics).
Send and receive lists enumerate the same global ids in ascending order on both sides, so no ids travel -- only rows.
On CUDA tensors everything but the collective itself is a kernel of libgp_b200.so working on preallocated staging
buffers (nothing allocates, nothing syncs: the exchange can be captured into the step's CUDA graph); host tensors
(gloo tests) use the same index lists with torch indexing."""
from __future__ import annotations

from typing import Dict

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
        send, recv = cat(lg.send), cat(lg.recv)
        self.send_idx, self.recv_idx = send.to(device), recv.to(device)
        self.n_send, self.n_recv = int(send.numel()), int(recv.numel())
        self.native = torch.device(device).type == "cuda"
        self._buf: Dict[tuple, torch.Tensor] = {}
        if self.native:
            self.send_i32, self.recv_i32 = self.send_idx.int(), self.recv_idx.int()
            # transpose maps: destination rows (unique owners' rows), CSR over the received list in ascending position
            order = torch.sort(send, stable=True).indices
            uniq, counts = torch.unique_consecutive(send[order], return_counts=True)
            self.add_rows = uniq.int().to(device)
            self.add_rowptr = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]).int().to(device)
            self.add_order = order.int().to(device)

    def _staging(self, name: str, rows: int, h: int, dtype, device) -> torch.Tensor:
        key = (name, rows, h, dtype)
        b = self._buf.get(key)
        if b is None:
            b = self._buf[key] = torch.zeros((rows, h), dtype=dtype, device=device)
        return b

    def exchange_(self, x: torch.Tensor, group=None) -> None:
        """x[recv_idx] <- rows x[send_idx] of the owning ranks (in place on the ghost rows of x)."""
        h = x.shape[1]
        if self.native and x.is_cuda:
            from .. import ops
            out = self._staging("fs", self.n_send, h, x.dtype, x.device)
            inp = self._staging("fr", self.n_recv, h, x.dtype, x.device)
            if self.n_send:
                ops.halo_pack(x, self.send_i32, out)
            dist.all_to_all_single(inp, out, self.recv_counts, self.send_counts, group=group)
            if self.n_recv:
                ops.halo_unpack(x, self.recv_i32, inp)
            return
        out = x.index_select(0, self.send_idx).contiguous()
        inp = torch.empty((self.n_recv, h), dtype=x.dtype, device=x.device)
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
        if self.native and dx.is_cuda:
            from .. import ops
            out = self._staging("bs", self.n_recv, h, dx.dtype, dx.device)
            inp = self._staging("br", self.n_send, h, dx.dtype, dx.device)
            zeros = self._staging("bz", self.n_recv, h, dx.dtype, dx.device)      # never written: stays zero
            if self.n_recv:
                ops.halo_pack(dx, self.recv_i32, out)
            dist.all_to_all_single(inp, out, self.send_counts, self.recv_counts, group=group)
            if self.n_send:
                ops.halo_unpack_add(dx, self.add_rows, self.add_rowptr, self.add_order, inp)
            if self.n_recv:
                ops.halo_unpack(dx, self.recv_i32, zeros)
            return
        out = dx.index_select(0, self.recv_idx).contiguous()
        inp = torch.empty((self.n_send, h), dtype=dx.dtype, device=dx.device)
        dist.all_to_all_single(inp, out, self.send_counts, self.recv_counts, group=group)
        dx.index_add_(0, self.send_idx, inp)           # a row sent to several peers collects all of them
        dx.index_fill_(0, self.recv_idx, 0.0)
