"""Graph construction on the device: the reference's preprocessing transforms with the same names and arguments
(graphphysics/dataset/preprocessing.py), running as kernels of libgp_b200.so on CUDA tensors instead of PyG / scipy on
the host -- the step right before the model, and for DeformingPlate (world edges follow the moving obstacle) one that
runs every roll-out frame.

    face_to_edge(graph)                 T.FaceToEdge(remove_faces=False)        preprocessing.py:410-424, torch_graph.py:194-210
    add_edge_features(graph)            T.Cartesian + T.Distance (norm=False)   preprocessing.py:16-23
    add_obstacles_next_pos(graph, ...)                                          preprocessing.py:47-89
    add_world_edges(graph, ...)         cKDTree.query_pairs + mask + to_undirected   preprocessing.py:92-140
    add_world_pos_features(graph, ...)                                          preprocessing.py:143-175
    add_noise(graph, ...)                                                       preprocessing.py:177-238
    build_preprocessing(...)            the same composition order              preprocessing.py:372-441

`graph` is a graphphysics_b200.graph.Data (or PyG Data) whose tensors live on a CUDA device; `graph.face` is (3, F) like
PyG, `graph.tetra` (4, T) is used when there are no faces.  Integer outputs are bit-exact against the CPU oracle; the
edge features are computed in fp32 without contraction and match it bit for bit as well.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import partial
from typing import Callable, List, Optional, Union

import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr
from .utils.nodetype import NodeType


def _cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"graphphysics_b200.preprocessing: {what} must be a CUDA tensor (there is no CPU fallback)")
    return t


def coalesce(cand_row: torch.Tensor, cand_col: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """Unique directed pairs sorted by (row, col) -> int64 (2, E) (PyG coalesce / to_undirected's last step)."""
    _cuda(cand_row, "candidate list")
    cand_row, cand_col = cand_row.long().contiguous(), cand_col.long().contiguous()
    n = cand_row.numel()
    L = lib()
    L.gp_coalesce_workspace_bytes.restype = C.c_int64
    ws = torch.empty(int(L.gp_coalesce_workspace_bytes(C.c_int64(n), C.c_int32(num_nodes))), dtype=torch.uint8, device=cand_row.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=cand_row.device)
    check(L.gp_coalesce_count(C.c_void_p(ptr(cand_row)), C.c_void_p(ptr(cand_col)), C.c_int64(n), C.c_int32(num_nodes), C.c_void_p(ptr(ws)),
                              C.c_void_p(ptr(cnt)), C.c_void_p(stream_ptr())), "gp_coalesce_count")
    ops._launched(13)
    E = int(cnt.item())                      # the one host round trip: the size of the output
    out = torch.empty((2, E), dtype=torch.int64, device=cand_row.device)
    check(L.gp_coalesce_write(C.c_void_p(ptr(cand_row)), C.c_void_p(ptr(cand_col)), C.c_int64(n), C.c_int32(num_nodes), C.c_void_p(ptr(ws)),
                              C.c_void_p(ptr(out[0])), C.c_void_p(ptr(out[1])), C.c_void_p(stream_ptr())), "gp_coalesce_write")
    ops._launched()
    return out


def cells_to_edge_index(cells: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """cells: (3, F) triangles or (4, T) tetrahedra (vertex-major, like PyG `face`) -> undirected coalesced edge_index."""
    _cuda(cells, "face / tetra")
    cells = cells.long().contiguous()
    verts, n = cells.shape
    per = 6 if verts == 3 else 12
    cand = torch.empty((2, n * per), dtype=torch.int64, device=cells.device)
    check(lib().gp_cell_edge_candidates(C.c_void_p(ptr(cells)), C.c_int64(n), C.c_int32(verts), C.c_int32(0), C.c_void_p(ptr(cand[0])),
                                        C.c_void_p(ptr(cand[1])), C.c_void_p(stream_ptr())), "gp_cell_edge_candidates")
    ops._launched()
    return coalesce(cand[0], cand[1], num_nodes)


def face_to_edge(graph):
    """T.FaceToEdge(remove_faces=False): graph.edge_index from graph.face (tetrahedra were already turned into their four
    triangles by the reference's mesh loader; passing graph.tetra directly gives the same edge set)."""
    cells = graph.face if graph.face is not None else graph.tetra
    if cells is None:
        raise ValueError("face_to_edge needs graph.face (3, F) or graph.tetra (4, T)")
    graph.edge_index = cells_to_edge_index(cells, graph.x.shape[0] if graph.x is not None else graph.pos.shape[0])
    return graph


def k_hop_edge_index(edge_index: torch.Tensor, num_hops: int, num_nodes: int) -> torch.Tensor:
    """compute_k_hop_edge_index (graphphysics/utils/torch_graph.py:14-54): the pattern of adj_k <- adj_k + adj_k . adj with
    self loops removed, (num_hops - 1) times, sorted by (row, col).  Per hop: gp_khop_candidates + the coalesce kernels."""
    _cuda(edge_index, "edge_index")
    ei = edge_index.long().contiguous()
    adj = coalesce(ei[0], ei[1], num_nodes)                          # sorted by (row, col), duplicates merged
    if num_hops <= 1:
        return adj
    deg = torch.bincount(adj[0], minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=ei.device)
    rowptr[1:] = torch.cumsum(deg, 0)
    adj_col = adj[1].contiguous()
    cur = adj
    for _ in range(num_hops - 1):
        cur = cur[:, cur[0] != cur[1]]                                # (a self loop in the input adjacency)
        rk, ck = cur[0].contiguous(), cur[1].contiguous()
        per = 1 + deg[ck]
        offsets = torch.cumsum(per, 0) - per
        total = int(per.sum().item())                                 # size of the candidate list: one host round trip per hop
        cand = torch.empty((2, total), dtype=torch.int64, device=ei.device)
        check(lib().gp_khop_candidates(C.c_void_p(ptr(rk)), C.c_void_p(ptr(ck)), C.c_int64(rk.numel()), C.c_void_p(ptr(rowptr)),
                                       C.c_void_p(ptr(adj_col)), C.c_void_p(ptr(offsets)), C.c_void_p(ptr(cand[0])), C.c_void_p(ptr(cand[1])),
                                       C.c_void_p(stream_ptr())), "gp_khop_candidates")
        ops._launched()
        cur = coalesce(cand[0], cand[1], num_nodes)
    return cur


def k_hop_graph(graph, num_hops: int, add_edge_features_to_khop: bool = False):
    """compute_k_hop_graph (torch_graph.py:57-105): the graph with its k-hop edge_index, optionally with fresh
    [pos_i - pos_j, distance] edge features."""
    if num_hops == 1:
        return graph
    n = graph.x.shape[0] if graph.x is not None else graph.pos.shape[0]
    out = graph.clone() if hasattr(graph, "clone") else graph
    out.edge_index = k_hop_edge_index(graph.edge_index, num_hops, n)
    if add_edge_features_to_khop:
        out.edge_attr = edge_features(out.pos, out.edge_index)
    return out


def edge_features(pos: torch.Tensor, edge_index: torch.Tensor, out: Optional[torch.Tensor] = None, col: int = 0) -> torch.Tensor:
    """[pos[row] - pos[col], ||pos[col] - pos[row]||] per edge (fp32, dim + 1 columns), optionally written into columns
    [col, col + dim + 1) of an existing [E, >=] buffer."""
    _cuda(pos, "pos")
    pos = pos.float().contiguous()
    ei = edge_index.long().contiguous()
    E, dim = ei.shape[1], pos.shape[1]
    if out is None:
        out = torch.empty((E, dim + 1), dtype=torch.float32, device=pos.device)
    view = out[:, col:]
    check(lib().gp_edge_features(C.c_void_p(ptr(pos)), C.c_int32(pos.stride(0)), C.c_int32(dim), C.c_void_p(ptr(ei[0])), C.c_void_p(ptr(ei[1])),
                                 C.c_int64(E), C.c_void_p(ptr(view)), C.c_int32(out.stride(0)), C.c_void_p(stream_ptr())), "gp_edge_features")
    ops._launched()
    return out


def add_edge_features(graph=None):
    """With no argument: the reference's list of transforms (preprocessing.py:16-23).  With a graph: applies them."""
    def _apply(g):
        new = edge_features(g.pos, g.edge_index)
        g.edge_attr = new if g.edge_attr is None else torch.cat([g.edge_attr.reshape(new.shape[0], -1), new], dim=1)
        return g
    return [_apply] if graph is None else _apply(graph)


def add_obstacles_next_pos(graph, world_pos_index_start: int, world_pos_index_end: int, node_type_index: int):
    """preprocessing.py:47-89 (row-wise bookkeeping on N x F node features; torch ops on the device)."""
    world_pos = graph.x[:, world_pos_index_start:world_pos_index_end]
    other = graph.x[:, world_pos_index_end:]
    disp = graph.y[:, world_pos_index_start:world_pos_index_end] - world_pos
    node_type = graph.x[:, node_type_index - 3]
    is_obs = node_type == NodeType.OBSTACLE
    mean_disp = disp[is_obs].mean(dim=0)
    disp = torch.where(is_obs[:, None], disp, mean_disp[None, :].expand_as(disp))
    graph.x = torch.cat([world_pos, disp, other], dim=1)
    return graph


def world_pairs(world_pos: torch.Tensor, node_type: torch.Tensor, radius: float):
    """Both directions of every OBSTACLE-NORMAL node pair within `radius` -> (row, col) int64 tensors."""
    _cuda(world_pos, "world_pos")
    N = world_pos.shape[0]
    wp = world_pos.float()
    if wp.shape[1] < 3:
        wp = torch.cat([wp, wp.new_zeros(N, 3 - wp.shape[1])], dim=1)
    wp = wp.contiguous()
    nt = node_type.float().contiguous()
    L = lib()
    L.gp_world_pairs_workspace_bytes.restype = C.c_int64
    ws = torch.empty(int(L.gp_world_pairs_workspace_bytes(C.c_int32(N))), dtype=torch.uint8, device=wp.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=wp.device)
    check(L.gp_world_pairs_count(C.c_void_p(ptr(wp)), C.c_int32(3), C.c_void_p(ptr(nt)), C.c_int32(1), C.c_int32(N), C.c_double(radius),
                                 C.c_int32(int(NodeType.NORMAL)), C.c_int32(int(NodeType.OBSTACLE)), C.c_void_p(ptr(ws)), C.c_void_p(ptr(cnt)),
                                 C.c_void_p(stream_ptr())), "gp_world_pairs_count")
    ops._launched(12)
    P = int(cnt.item())
    pairs = torch.empty((2, 2 * P), dtype=torch.int64, device=wp.device)
    if P:
        check(L.gp_world_pairs_fill(C.c_void_p(ptr(wp)), C.c_int32(3), C.c_void_p(ptr(nt)), C.c_int32(1), C.c_int32(N), C.c_double(radius),
                                    C.c_int32(int(NodeType.OBSTACLE)), C.c_void_p(ptr(ws)), C.c_void_p(ptr(pairs[0])), C.c_void_p(ptr(pairs[1])),
                                    C.c_void_p(stream_ptr())), "gp_world_pairs_fill")
        ops._launched()
    return pairs[0], pairs[1]


def add_world_edges(graph, world_pos_index_start: int, world_pos_index_end: int, node_type_index: int, radius: float = 0.03):
    """preprocessing.py:92-140: radius search in world space between OBSTACLE and NORMAL nodes, merged with the mesh edges
    and made undirected / coalesced."""
    world_pos = graph.x[:, world_pos_index_start:world_pos_index_end]
    r, c = world_pairs(world_pos, graph.x[:, node_type_index], radius)
    ei = graph.edge_index.long()
    graph.edge_index = coalesce(torch.cat([r, ei[0], ei[1]]), torch.cat([c, ei[1], ei[0]]), graph.x.shape[0])
    return graph


def add_world_pos_features(graph, world_pos_index_start: int, world_pos_index_end: int):
    """preprocessing.py:143-175: edge_attr += [wp[senders] - wp[receivers], ||.||]."""
    world_pos = graph.x[:, world_pos_index_start:world_pos_index_end]
    E, dim = graph.edge_index.shape[1], world_pos.shape[1]
    old = graph.edge_attr.reshape(E, -1).float()
    out = torch.empty((E, old.shape[1] + dim + 1), dtype=torch.float32, device=old.device)
    out[:, : old.shape[1]] = old
    graph.edge_attr = edge_features(world_pos, graph.edge_index, out=out, col=old.shape[1])
    return graph


def apply_noise_(x: torch.Tensor, noise: torch.Tensor, start: int, end: int, scale: float, node_type_index: int) -> None:
    """x[:, start:end] += noise * scale on NORMAL rows, in place (gp_add_noise)."""
    check(lib().gp_add_noise(C.c_void_p(ptr(x)), C.c_int32(x.stride(0)), C.c_int32(x.shape[0]), C.c_int32(start), C.c_int32(end),
                             C.c_int32(node_type_index), C.c_int32(int(NodeType.NORMAL)), C.c_void_p(ptr(noise.contiguous())), C.c_float(scale),
                             C.c_void_p(stream_ptr())), "gp_add_noise")
    ops._launched()


def add_noise(graph, noise_index_start: Union[int, List[int]], noise_index_end: Union[int, List[int]],
              noise_scale: Union[float, List[float]], node_type_index: int, t: Optional[float] = None,
              generator: Optional[torch.Generator] = None):
    """preprocessing.py:177-238: Gaussian noise on the given feature columns of NORMAL nodes (in place on graph.x), with the
    curriculum scale 10 * std * (1 + cos(pi t)) when `t` is given.  The draw is torch's device generator (works under
    CUDA-graph capture); the masked scale-and-add is gp_add_noise."""
    if isinstance(noise_index_start, int):
        noise_index_start = [noise_index_start]
    if isinstance(noise_index_end, int):
        noise_index_end = [noise_index_end]
    if isinstance(noise_scale, float):
        noise_scale = [noise_scale] * len(noise_index_start)
    if len(noise_index_start) != len(noise_index_end):
        raise ValueError("noise_index_start and noise_index_end must have the same length.")
    if len(noise_scale) != len(noise_index_start):
        raise ValueError("noise_scale must have the same length as noise_index_start and noise_index_end.")
    x = _cuda(graph.x, "graph.x")
    if x.dtype != torch.float32 or x.stride(1) != 1:
        raise ValueError("add_noise works in place on an fp32 graph.x with contiguous rows")
    for start, end, scale in zip(noise_index_start, noise_index_end, noise_scale):
        scale_ = 10 * scale * (1 + math.cos(t * math.pi)) if t is not None else scale
        noise = torch.randn((x.shape[0], end - start), dtype=torch.float32, device=x.device, generator=generator)
        apply_noise_(x, noise, start, end, scale_, node_type_index)
    return graph


class Compose:
    """torch_geometric.transforms.Compose: applies the transforms in order."""

    def __init__(self, transforms: List[Callable]):
        self.transforms = transforms

    def __call__(self, graph):
        for t in self.transforms:
            graph = t(graph)
        return graph


def build_preprocessing(noise_parameters: Optional[dict] = None, world_pos_parameters: Optional[dict] = None,
                        add_edges_features: bool = True, extra_node_features=None, extra_edge_features=None) -> Compose:
    """The reference's pipeline in the reference's order (preprocessing.py:372-441)."""
    pre: List[Callable] = []
    if extra_node_features is not None:
        pre.extend(extra_node_features if isinstance(extra_node_features, list) else [extra_node_features])
    if world_pos_parameters is not None:
        w = world_pos_parameters
        pre.extend([
            partial(add_obstacles_next_pos, world_pos_index_start=w["world_pos_index_start"], world_pos_index_end=w["world_pos_index_end"],
                    node_type_index=w["node_type_index"]),
            face_to_edge,
            partial(add_world_edges, world_pos_index_start=w["world_pos_index_start"], world_pos_index_end=w["world_pos_index_end"],
                    node_type_index=w["node_type_index"], radius=w.get("radius", 0.03)),
        ])
        pre.extend(add_edge_features())
    else:
        pre.append(face_to_edge)
        if add_edges_features:
            pre.extend(add_edge_features())
    if noise_parameters is not None:
        n = noise_parameters
        pre.insert(1, partial(add_noise, noise_index_start=n["noise_index_start"], noise_index_end=n["noise_index_end"],
                              noise_scale=n["noise_scale"], node_type_index=n["node_type_index"]))
    if extra_edge_features is not None:
        pre.extend(extra_edge_features if isinstance(extra_edge_features, list) else [extra_edge_features])
    return Compose(pre)
