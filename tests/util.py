"""Shared helpers for the parity tests."""
import numpy as np
import torch


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """max |got-ref| / max |ref|  (norm-relative, robust to near-zero entries)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def l2_rel(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(x.dtype)


def random_sorted_graph(num_nodes: int, num_edges: int, seed: int = 0, max_degree_node: int = -1):
    """Random directed graph, edges sorted by receiver; some nodes have no in-edges and
    (optionally) one node gets a very long segment."""
    rng = np.random.default_rng(seed)
    dst = rng.integers(0, num_nodes, num_edges)
    if max_degree_node >= 0:
        dst[: num_edges // 4] = max_degree_node
    dst = np.sort(dst)
    src = rng.integers(0, num_nodes, num_edges)
    return torch.from_numpy(src.astype(np.int64)), torch.from_numpy(dst.astype(np.int64))
