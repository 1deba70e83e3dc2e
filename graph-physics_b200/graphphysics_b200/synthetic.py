"""Synthetic meshes and batches of the shapes BASELINE.json names (there is no dataset access):
triangulated channel with a cylinder hole (CylinderFlow-like) and a tetrahedral box.  Edge
construction follows what the reference's preprocessing does to a mesh
(graphphysics/dataset/preprocessing.py:16-23, 410-424; graphphysics/utils/torch_graph.py:194-210):
FaceToEdge -> undirected coalesced edges sorted by (sender, receiver), then
edge_attr = [pos[row]-pos[col], ||pos[col]-pos[row]||]."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .graph import Data
from .utils.nodetype import NodeType


def channel_mesh(nx: int = 64, ny: int = 32, jitter: float = 0.3, seed: int = 0,
                 hole: Optional[Tuple[float, float, float]] = (0.33, 0.2, 0.06)):
    """pos (N,2) float32, triangles (T,3) int64 of a 1.6 x 0.41 channel with a circular hole."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.linspace(0, 1.6, nx), np.linspace(0, 0.41, ny), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel()], -1)
    idx = np.arange(nx * ny).reshape(nx, ny)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    tris = np.concatenate([np.stack([a, b, c], -1), np.stack([b, d, c], -1)])
    if jitter:
        inner = np.ones((nx, ny), bool)
        inner[0], inner[-1], inner[:, 0], inner[:, -1] = False, False, False, False
        pos = pos + inner.ravel()[:, None] * rng.uniform(-jitter, jitter, pos.shape) * np.array([1.6 / nx, 0.41 / ny])
    if hole is not None:
        cx, cy, r = hole
        keep = ((pos[:, 0] - cx) ** 2 + (pos[:, 1] - cy) ** 2) > r * r
        tris = tris[keep[tris].all(1)]
        used = np.zeros(len(pos), bool)
        used[tris.ravel()] = True
        remap = np.cumsum(used) - 1
        pos, tris = pos[used], remap[tris]
    return pos.astype(np.float32), tris.astype(np.int64)


def box_tet_mesh(nx: int, ny: int, nz: int):
    """pos (N,3), tetrahedra (T,4): structured box, 6 tets per cell."""
    xs, ys, zs = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny), np.linspace(0, 1, nz), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], -1)
    idx = np.arange(nx * ny * nz).reshape(nx, ny, nz)
    v = [idx[i:nx - 1 + i, j:ny - 1 + j, k:nz - 1 + k].ravel() for i in (0, 1) for j in (0, 1) for k in (0, 1)]
    v000, v001, v010, v011, v100, v101, v110, v111 = v
    tets = np.concatenate([np.stack(t, -1) for t in (
        (v000, v100, v110, v111), (v000, v110, v010, v111), (v000, v010, v011, v111),
        (v000, v011, v001, v111), (v000, v001, v101, v111), (v000, v101, v100, v111))])
    return pos.astype(np.float32), tets.astype(np.int64)


def faces_of_cells(cells: np.ndarray) -> np.ndarray:
    """Triangles stay; a tetrahedron contributes its 4 triangles (torch_graph.py:194-210)."""
    if cells.shape[1] == 3:
        return cells
    t = cells
    return np.concatenate([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]], axis=0)


def mesh_edges(faces: np.ndarray, num_nodes: int) -> np.ndarray:
    """int64 (2,E): undirected, coalesced, sorted by (row, col) -- PyG FaceToEdge semantics."""
    f = np.asarray(faces, dtype=np.int64).T
    ei = np.concatenate([f[:2], f[1:], f[::2]], axis=1)
    both = np.concatenate([ei, ei[::-1]], axis=1)
    key = np.unique(both[0] * np.int64(num_nodes) + both[1])
    return np.stack([key // num_nodes, key % num_nodes])


def mesh_edge_attr(pos: np.ndarray, edge_index: np.ndarray) -> np.ndarray:
    row, col = edge_index
    cart = pos[row] - pos[col]
    return np.concatenate([cart, np.linalg.norm(pos[col] - pos[row], axis=-1, keepdims=True)], -1).astype(np.float32)


def cylinder_node_types(pos: np.ndarray, hole=(0.33, 0.2, 0.06)) -> np.ndarray:
    """NORMAL inside; INFLOW left, OUTFLOW right, WALL top/bottom and around the cylinder."""
    t = np.full(len(pos), int(NodeType.NORMAL), np.int64)
    eps = 1e-6
    t[(pos[:, 1] < eps) | (pos[:, 1] > 0.41 - eps)] = int(NodeType.WALL_BOUNDARY)
    if hole is not None:
        cx, cy, r = hole
        d = np.sqrt((pos[:, 0] - cx) ** 2 + (pos[:, 1] - cy) ** 2)
        t[d < r + 0.035] = int(NodeType.WALL_BOUNDARY)
    t[pos[:, 0] < eps] = int(NodeType.INFLOW)
    t[pos[:, 0] > 1.6 - eps] = int(NodeType.OUTFLOW)
    return t


def cylinder_flow_batch(batch_size: int = 32, nx: int = 64, ny: int = 32, seed: int = 0, pin: bool = False) -> Data:
    """Union graph of `batch_size` CylinderFlow-shaped meshes (PyG collate: features concatenated,
    edge_index offset per graph).  Raw node columns follow the H5 layout [vx, vy, node_type, time]
    (graphphysics/utils/hierarchical.py:121-126); y is the next-step velocity.  CPU tensors."""
    rng = np.random.default_rng(seed)
    xs, ys, eis, eas, poss = [], [], [], [], []
    off = 0
    for b in range(batch_size):
        pos, tris = channel_mesh(nx, ny, seed=seed * 1000 + b)
        ei = mesh_edges(tris, len(pos))
        nt = cylinder_node_types(pos)
        vel = rng.standard_normal((len(pos), 2)).astype(np.float32)
        x = np.concatenate([vel, nt[:, None].astype(np.float32), np.zeros((len(pos), 1), np.float32)], 1)
        y = vel + 0.1 * rng.standard_normal((len(pos), 2)).astype(np.float32)
        xs.append(x); ys.append(y); poss.append(pos)
        eis.append(ei + off); eas.append(mesh_edge_attr(pos, ei))
        off += len(pos)
    t = dict(x=torch.from_numpy(np.concatenate(xs)), y=torch.from_numpy(np.concatenate(ys)),
             pos=torch.from_numpy(np.concatenate(poss)), edge_index=torch.from_numpy(np.concatenate(eis, 1)),
             edge_attr=torch.from_numpy(np.concatenate(eas)))
    if pin and torch.cuda.is_available():
        t = {k: v.pin_memory() for k, v in t.items()}
    return Data(**t)


def kuhn_box_graph(nx: int, ny: int, nz: int):
    """pos (N,3) float32 and the directed edge list (2,E) int64 of the 6-tetrahedra-per-cell (Kuhn)
    triangulation of a structured box, built from its 7 edge directions instead of from the
    tetrahedra (same graph as mesh_edges(faces_of_cells(box_tet_mesh(...))), without the 100M-key
    sort): 14 neighbours per interior node.  Sorted by (row, col) like PyG's coalesce."""
    xs, ys, zs = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny), np.linspace(0, 1, nz), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], -1).astype(np.float32)
    idx = np.arange(nx * ny * nz, dtype=np.int64).reshape(nx, ny, nz)
    rows, cols = [], []
    for dx, dy, dz in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)):
        a = idx[: nx - dx, : ny - dy, : nz - dz].ravel()
        b = idx[dx:, dy:, dz:].ravel()
        rows += [a, b]
        cols += [b, a]
    row, col = np.concatenate(rows), np.concatenate(cols)
    order = np.argsort(row * np.int64(nx * ny * nz) + col, kind="stable")
    return pos, np.stack([row[order], col[order]])


def deforming_plate_sample(nx: int = 21, ny: int = 21, nz: int = 3, seed: int = 0):
    """DeformingPlate-shaped sample (BASELINE.json configs[2]; no dataset access): a thin tetrahedral plate (~1.3k NORMAL
    nodes, HANDLE nodes along one edge) and a small rigid OBSTACLE block hovering within the world-edge radius above it.
    Returns numpy arrays: mesh_pos (N,3) fp32, tetra (T,4) int64, x_raw (N,4) = [world_pos, node_type], y (N,3) = the next
    world_pos (the obstacle moves down, the plate gives way a little)."""
    rng = np.random.default_rng(seed)
    p, t = box_tet_mesh(nx, ny, nz)
    p = p * np.array([0.5, 0.5, 0.02], np.float32)
    o, to = box_tet_mesh(5, 5, 3)
    o = o * np.array([0.08, 0.08, 0.04], np.float32) + np.array([0.21, 0.21, 0.03], np.float32)
    pos = np.concatenate([p, o]).astype(np.float32)
    tets = np.concatenate([t, to + len(p)])
    ntype = np.full(len(pos), int(NodeType.NORMAL), np.int64)
    ntype[: len(p)][p[:, 0] < 1e-6] = int(NodeType.HANDLE)
    ntype[len(p):] = int(NodeType.OBSTACLE)
    world = pos + np.concatenate([0.002 * rng.standard_normal(p.shape), np.zeros_like(o)]).astype(np.float32)
    nxt = world.copy()
    nxt[len(p):, 2] -= 0.002
    nxt[: len(p)] += (0.0005 * rng.standard_normal(p.shape)).astype(np.float32)
    x_raw = np.concatenate([world, ntype[:, None].astype(np.float32)], 1).astype(np.float32)
    return pos, tets, x_raw, nxt.astype(np.float32)
