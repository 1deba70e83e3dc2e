"""Data-parallel glue: one process per GPU, batches of independent graphs per rank.

The reference is single-device (`devices: 1`, graphphysics/train.py:276-279); the equivalent of its
result on N GPUs is "one device with the N-times larger batch", so after the backward the flat
gradient buffer is averaged over ranks with ONE all-reduce (NCCL on GPUs, gloo in the CPU tests), and
the online normalisers (graphphysics/models/layers.py:331-377) accumulate the statistics of the
global batch so every rank normalises identically.  Works on any backend / device."""
from __future__ import annotations

import torch
import torch.distributed as dist


def broadcast_(flat: torch.Tensor, group=None, src: int = 0) -> None:
    dist.broadcast(flat, src=src, group=group)


def allreduce_mean_(flat_grad: torch.Tensor, group=None) -> None:
    """flat_grad <- mean over ranks (sum all-reduce, then scale)."""
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    flat_grad.mul_(1.0 / dist.get_world_size(group))


def _local_stats(norm, data: torch.Tensor):
    """[sum, sum of squares, count] of the local rows, or None when the normaliser is frozen / absent."""
    if norm is None:
        return None
    if norm._host_calls is None:
        norm._host_calls = int(norm._num_accumulations.item())
    if norm._host_calls >= norm._max_accumulations:
        return None
    d = data.detach()
    if norm._native(d):
        from .. import ops
        return ops.normalizer_update(ops.normalizer_stats(d), d.shape[0], want_stats=True)
    return torch.cat([d.sum(0), (d ** 2).sum(0), torch.full((1,), float(d.shape[0]), dtype=d.dtype, device=d.device)])   # (no H2D copy: capturable)


def _apply_stats(norm, stats: torch.Tensor) -> None:
    """Normalizer._accumulate (layers.py:363-377) with the sums taken over all ranks."""
    k = (stats.numel() - 1) // 2
    if stats.is_cuda and norm._acc_sum.is_cuda:
        from .. import ops
        ops.normalizer_accumulate(stats.contiguous(), norm)
        norm._host_calls += 1
        return
    gate = (norm._num_accumulations < norm._max_accumulations).to(torch.float32)     # device-side freeze (graph replay)
    norm._acc_sum += gate * stats[:k][None]
    norm._acc_sum_squared += gate * stats[k:2 * k][None]
    norm._acc_count += gate * stats[2 * k]
    norm._num_accumulations += gate
    norm._host_calls += 1


def accumulate_normalizers_globally(sim, batch, group=None) -> None:
    """What Simulator._build_input_graph(is_training=True) accumulates (simulator.py:145-167), over the
    global batch: the statistics of all three normalisers travel in ONE all-reduce.  Call before building the
    input graph with accumulate=False."""
    pre = batch.x[:, sim.output_index_start:sim.output_index_end]
    items = [(sim._output_normalizer, batch.y - pre)]
    if sim._node_normalizer is not None:
        items.append((sim._node_normalizer, sim.node_features(batch)))
    if sim._edge_normalizer is not None:
        items.append((sim._edge_normalizer, batch.edge_attr))
    live = [(n, st) for n, st in ((n, _local_stats(n, d)) for n, d in items) if st is not None]
    if not live:
        return
    buf = torch.cat([st for _, st in live])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for n, st in live:
        _apply_stats(n, buf[off:off + st.numel()])
        off += st.numel()
