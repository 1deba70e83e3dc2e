"""GPU: graph construction kernels (csrc/mesh_ops.cu, graphphysics_b200/preprocessing.py) against the CPU oracle, BIT-EXACT:
FaceToEdge on the reference's own fixture meshes (11 070 / 291 144 directed edges, the counts the reference's tests pin:
tests/graphphysics/dataset/test_xdmfdataset.py:31,46), Cartesian + Distance edge features, world edges (cKDTree radius
search + node-type mask + to_undirected), world-position features, noise injection, and the whole build_preprocessing
pipeline on a DeformingPlate-shaped sample (BASELINE.json configs[2])."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_face_to_edge_cylinder_and_aneurysm_bit_exact():
    from oracle import gp_oracle as O
    from graphphysics_b200 import preprocessing as P
    from graphphysics_b200.graph import Data
    c = np.load(os.path.join(G, "cylinder_mesh.npz"))
    ref = O.face_to_edge(c["triangles"], 1923)
    g = P.face_to_edge(Data(x=torch.zeros(1923, 1, device=DEV), face=_t(c["triangles"].T.astype(np.int64))))
    assert g.edge_index.dtype == torch.int64 and tuple(g.edge_index.shape) == (2, 11070)
    assert np.array_equal(g.edge_index.cpu().numpy(), ref)
    a = np.load(os.path.join(G, "aneurysm_mesh.npz"))
    ref = O.face_to_edge(O.tetra_to_faces(a["tets"]), 22535)
    g = P.face_to_edge(Data(x=torch.zeros(22535, 1, device=DEV), tetra=_t(a["tets"].T.astype(np.int64))))
    assert tuple(g.edge_index.shape) == (2, 291144)
    assert np.array_equal(g.edge_index.cpu().numpy(), ref)
    # the reference's loader hands FaceToEdge the four triangles of every tetrahedron (torch_graph.py:194-210): same result
    cells = torch.from_numpy(a["tets"].T.astype(np.int64))
    face = torch.cat([cells[0:3], cells[1:4], torch.stack([cells[2], cells[3], cells[0]]), torch.stack([cells[3], cells[0], cells[1]])], dim=1)
    g2 = P.face_to_edge(Data(x=torch.zeros(22535, 1, device=DEV), face=face.to(DEV)))
    assert torch.equal(g2.edge_index, g.edge_index)


def test_coalesce_edge_cases():
    from graphphysics_b200 import preprocessing as P
    # duplicates, self loops, isolated nodes, one hub row with a long bucket, empty input
    rng = np.random.default_rng(0)
    n = 500
    row = np.concatenate([rng.integers(0, n, 4000), np.full(3000, 7), np.arange(50), np.arange(50)])
    col = np.concatenate([rng.integers(0, n, 4000), rng.integers(0, n, 3000), np.arange(50), np.arange(50)])
    row[row == 11] = 12                                            # node 11 has no outgoing entries
    key = np.unique(row.astype(np.int64) * n + col)
    out = P.coalesce(_t(row.astype(np.int64)), _t(col.astype(np.int64)), n).cpu().numpy()
    assert np.array_equal(out, np.stack([key // n, key % n]))
    empty = P.coalesce(torch.zeros(0, dtype=torch.int64, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), 5)
    assert tuple(empty.shape) == (2, 0)


@pytest.mark.parametrize("dim", [2, 3])
def test_edge_features_bit_exact(dim):
    from oracle import gp_oracle as O
    from graphphysics_b200 import preprocessing as P
    c = np.load(os.path.join(G, "cylinder_mesh.npz"))
    pos = np.ascontiguousarray(c["points"][:, :dim])
    if dim == 3:
        pos = pos + np.random.default_rng(1).standard_normal(pos.shape).astype(np.float32) * 0.01
    ei = O.face_to_edge(c["triangles"], 1923)
    ref = O.edge_features(pos, ei).astype(np.float32)
    got = P.edge_features(_t(pos), _t(ei)).cpu().numpy()
    assert got.shape == (11070, dim + 1)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def _plate():
    from graphphysics_b200.synthetic import deforming_plate_sample
    return deforming_plate_sample(seed=3)


def test_world_edges_bit_exact():
    from oracle import gp_oracle as O
    from graphphysics_b200 import preprocessing as P
    from graphphysics_b200.graph import Data
    pos, tets, x_raw, y = _plate()
    n = len(pos)
    mesh_ei = O.face_to_edge(O.tetra_to_faces(tets), n)
    ref = O.world_edges(mesh_ei, x_raw[:, :3], x_raw[:, 3].astype(np.int64), n, 0.03)
    assert ref.shape[1] > mesh_ei.shape[1] + 100                  # the sample does have world edges
    g = Data(x=_t(x_raw), edge_index=_t(mesh_ei))
    g = P.add_world_edges(g, 0, 3, 3, radius=0.03)
    assert np.array_equal(g.edge_index.cpu().numpy(), ref)
    # a lattice where many pairs sit EXACTLY at the radius (query_pairs is inclusive), and a cloud spread over few cells
    k = np.arange(6)
    lat = (np.stack(np.meshgrid(k, k, k, indexing="ij"), -1).reshape(-1, 3) * 0.25).astype(np.float32)
    types = (np.arange(len(lat)) % 2).astype(np.int64)            # alternate NORMAL / OBSTACLE
    none = np.zeros((2, 0), np.int64)
    ref = O.world_edges(none, lat, types, len(lat), 0.25)
    assert ref.shape[1] > 0
    g = P.add_world_edges(Data(x=_t(np.concatenate([lat, types[:, None].astype(np.float32)], 1)), edge_index=_t(none)), 0, 3, 3, radius=0.25)
    assert np.array_equal(g.edge_index.cpu().numpy(), ref)
    rng = np.random.default_rng(5)
    cloud = (rng.random((4000, 3)) * np.array([40.0, 1.0, 0.2])).astype(np.float32)
    types = rng.integers(0, 3, 4000).astype(np.int64)             # NORMAL / OBSTACLE / AIRFOIL (ignored)
    ref = O.world_edges(none, cloud, types, 4000, 0.11)
    g = P.add_world_edges(Data(x=_t(np.concatenate([cloud, types[:, None].astype(np.float32)], 1)), edge_index=_t(none)), 0, 3, 3, radius=0.11)
    assert np.array_equal(g.edge_index.cpu().numpy(), ref)


def test_noise_and_world_pos_features_bit_exact():
    from oracle import gp_oracle as O
    from graphphysics_b200 import preprocessing as P
    from graphphysics_b200.graph import Data
    pos, tets, x_raw, y = _plate()
    n = len(pos)
    noise = np.random.default_rng(2).standard_normal((n, 3)).astype(np.float32)
    for t in (None, 0.3):
        ref = O.add_noise(x_raw, noise, 0, 3, 0.003, 3, t=t)
        x = _t(x_raw.copy())
        import math
        scale = 10 * 0.003 * (1 + math.cos(t * math.pi)) if t is not None else 0.003
        P.apply_noise_(x, _t(noise), 0, 3, scale, 3)
        assert np.array_equal(x.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    # the transform itself: only NORMAL rows move, the node-type column is untouched, the draw has the right scale
    g = P.add_noise(Data(x=_t(x_raw.copy())), [0], [3], 0.003, 3, generator=torch.Generator(device=DEV).manual_seed(0))
    d = g.x.cpu().numpy() - x_raw
    normal = x_raw[:, 3] == 0
    assert np.all(d[~normal] == 0) and np.all(d[:, 3] == 0) and 0.002 < d[normal, :3].std() < 0.004
    ei = O.face_to_edge(O.tetra_to_faces(tets), n)
    ea = O.edge_features(pos, ei).astype(np.float32)
    ref = O.world_pos_features(ea, x_raw[:, :3], ei)
    g = P.add_world_pos_features(Data(x=_t(x_raw), edge_index=_t(ei), edge_attr=_t(ea)), 0, 3)
    assert np.array_equal(g.edge_attr.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_build_preprocessing_plate_pipeline():
    """plate.json's pipeline (world_pos_parameters on): obstacle displacement features, FaceToEdge, world edges, edge
    features -- every stage equal to the oracle's, and the result feeds the model."""
    from oracle import gp_oracle as O
    from graphphysics_b200 import preprocessing as P
    from graphphysics_b200.graph import Data
    pos, tets, x_raw, y = _plate()
    n = len(pos)
    pipe = P.build_preprocessing(world_pos_parameters={"world_pos_index_start": 0, "world_pos_index_end": 3, "node_type_index": 6})
    g = pipe(Data(x=_t(x_raw), y=_t(y), pos=_t(pos), tetra=_t(tets.T.astype(np.int64))))
    # add_obstacles_next_pos (preprocessing.py:47-89)
    disp = y - x_raw[:, :3]
    obs = x_raw[:, 3] == 1
    disp[~obs] = disp[obs].mean(0)
    assert np.allclose(g.x.cpu().numpy(), np.concatenate([x_raw[:, :3], disp, x_raw[:, 3:]], 1), atol=1e-7)
    ref_ei = O.world_edges(O.face_to_edge(O.tetra_to_faces(tets), n), x_raw[:, :3], x_raw[:, 3].astype(np.int64), n, 0.03)
    assert np.array_equal(g.edge_index.cpu().numpy(), ref_ei)
    ref_ea = O.edge_features(pos, ref_ei).astype(np.float32)
    assert np.array_equal(g.edge_attr.cpu().numpy().view(np.uint32), ref_ea.view(np.uint32))
    # one model step on the constructed graph (plate.json is the transformer; config 3 also names the epd variant)
    from graphphysics_b200.models.processors import EncodeProcessDecode, EncodeTransformDecode
    torch.manual_seed(0)
    feat = torch.randn(n, 6 + 9, device=DEV)
    out = EncodeTransformDecode(2, 15, 3, hidden_size=64, num_heads=4).to(DEV)(Data(x=feat, edge_index=g.edge_index))
    assert tuple(out.shape) == (n, 3) and torch.isfinite(out).all()
    out = EncodeProcessDecode(2, 15, 4, 3, hidden_size=128).to(DEV)(Data(x=feat, edge_index=g.edge_index, edge_attr=g.edge_attr))
    assert tuple(out.shape) == (n, 3) and torch.isfinite(out).all()


def test_xdmf_archive_to_training_step():
    """The reference's own XDMF test archive -> graph on the device -> one training step: I/O (graphphysics_b200.io, no meshio /
    h5py), graph construction (FaceToEdge + edge features kernels) and the fused model, end to end."""
    from graphphysics_b200 import preprocessing as P
    from graphphysics_b200.io.xdmf import XDMFTrajectory
    from graphphysics_b200.training.loop import Trainer
    meta = {"dt": 0.01, "features": {"velocity_x": {"type": "dynamic", "dtype": "float32"}, "velocity_y": {"type": "dynamic", "dtype": "float32"}}}
    traj = XDMFTrajectory(os.path.join(G, "mock_xdmf", "mock.xdmf"), meta, targets=["velocity_x", "velocity_y"])
    g = traj[0].to(DEV)                                   # x = [vx, vy, time = 0]: column 2 doubles as an all-NORMAL node type
    g.pos = g.pos[:, :2].contiguous()
    g = P.build_preprocessing()(g)
    assert tuple(g.edge_index.shape) == (2, 11070) and tuple(g.edge_attr.shape) == (11070, 3)
    cfg = {"model": {"type": "epd", "message_passing_num": 2, "hidden_size": 128, "node_input_size": 2, "output_size": 2, "edge_input_size": 3},
           "index": {"feature_index_start": 0, "feature_index_end": 2, "output_index_start": 0, "output_index_end": 2, "node_type_index": 2}}
    tr = Trainer(cfg, learning_rate=1e-3, num_steps=100, warmup=2, device=torch.device(DEV), seed=0)
    losses = [float(tr.training_step(g)) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_graph_construction_edge_cases():
    """One triangle, one tetrahedron, a mesh with unused nodes, no obstacle nodes (no world edges), radius below every distance."""
    from oracle import gp_oracle as O
    from graphphysics_b200 import preprocessing as P
    from graphphysics_b200.graph import Data
    tri = np.array([[0, 1, 2]], np.int64)
    g = P.face_to_edge(Data(x=torch.zeros(3, 1, device=DEV), face=_t(tri.T)))
    assert np.array_equal(g.edge_index.cpu().numpy(), O.face_to_edge(tri, 3)) and g.edge_index.shape[1] == 6
    tet = np.array([[3, 1, 0, 2]], np.int64)
    g = P.face_to_edge(Data(x=torch.zeros(6, 1, device=DEV), tetra=_t(tet.T)))          # nodes 4, 5 unused
    assert np.array_equal(g.edge_index.cpu().numpy(), O.face_to_edge(O.tetra_to_faces(tet), 6)) and g.edge_index.shape[1] == 12
    pos = np.random.default_rng(0).random((6, 3)).astype(np.float32)
    ea = P.edge_features(_t(pos), g.edge_index).cpu().numpy()
    assert np.array_equal(ea.view(np.uint32), O.edge_features(pos, g.edge_index.cpu().numpy()).astype(np.float32).view(np.uint32))
    x = np.concatenate([pos, np.zeros((6, 1), np.float32)], 1)                           # all NORMAL: no world edges
    g2 = P.add_world_edges(Data(x=_t(x), edge_index=g.edge_index.clone()), 0, 3, 3, radius=10.0)
    assert torch.equal(g2.edge_index, g.edge_index)
    x[0, 3] = 1.0                                                                        # one obstacle, radius smaller than any distance
    g3 = P.add_world_edges(Data(x=_t(x), edge_index=g.edge_index.clone()), 0, 3, 3, radius=1e-6)
    assert torch.equal(g3.edge_index, g.edge_index)
    g4 = P.add_world_edges(Data(x=_t(x), edge_index=g.edge_index.clone()), 0, 3, 3, radius=10.0)
    ref = O.world_edges(g.edge_index.cpu().numpy(), pos, x[:, 3].astype(np.int64), 6, 10.0)
    assert np.array_equal(g4.edge_index.cpu().numpy(), ref) and g4.edge_index.shape[1] > 12


@pytest.mark.parametrize("hops", [1, 2, 3])
def test_k_hop_edges_bit_exact(hops):
    """compute_k_hop_edge_index (graphphysics/utils/torch_graph.py:14-54) on the device: bit-exact against the oracle's sparse
    matrix powers; 32 638 edges at two hops on the reference's cylinder mesh (tests/graphphysics/dataset/test_xdmfdataset.py)."""
    from graphphysics_b200.preprocessing import k_hop_edge_index, k_hop_graph
    from graphphysics_b200.graph import Data
    from oracle import gp_oracle as O
    c = np.load(os.path.join(G, "cylinder_mesh.npz"))
    ei = O.face_to_edge(c["triangles"], 1923)
    got = k_hop_edge_index(_t(ei), hops, 1923)
    ref = O.khop_edges(ei, 1923, hops)
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), ref)
    if hops == 2:
        assert got.shape[1] == 32638
        g = k_hop_graph(Data(x=torch.zeros(1923, 1, device="cuda"), pos=_t(c["points"][:, :2].astype(np.float32)), edge_index=_t(ei)), 2, True)
        assert tuple(g.edge_attr.shape) == (32638, 3)
        np.testing.assert_array_equal(g.edge_attr.cpu().numpy(), O.edge_features(c["points"][:, :2].astype(np.float32), ref))


def test_k_hop_edge_cases():
    from graphphysics_b200.preprocessing import k_hop_edge_index
    from oracle import gp_oracle as O
    # a path 0-1-2-3 plus an isolated node, duplicate entries and a self loop in the input
    ei = np.array([[0, 1, 1, 2, 2, 3, 1, 2], [1, 0, 2, 1, 3, 2, 2, 2]])
    for hops in (1, 2, 3, 4):
        got = k_hop_edge_index(_t(ei), hops, 5).cpu().numpy()
        ref = O.khop_edges(ei, 5, hops) if hops > 1 else np.unique(ei[0] * 5 + ei[1])
        if hops == 1:
            assert np.array_equal(got[0] * 5 + got[1], ref)          # coalesced input, self loop kept (the reference's adj.indices())
        else:
            assert np.array_equal(got, ref), hops
    empty = k_hop_edge_index(torch.zeros((2, 0), dtype=torch.int64, device="cuda"), 3, 4)
    assert tuple(empty.shape) == (2, 0)
