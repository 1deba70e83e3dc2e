"""CPU: HDF5 / XDMF I/O without h5py / meshio (graphphysics_b200/io).

Fixtures: tests/golden/mock_xdmf/{mock.h5, mock.xdmf} are the reference's own test archive (reference tests/mock_xdmf/, written
by meshio + h5py: a DATA file, the ground truth for the reader); its content must equal the VTU-derived
tests/golden/cylinder_mesh.npz (same mesh, same velocity frames; 1923 nodes / 3612 triangles / 11 070 directed edges, the counts
the reference's tests pin: tests/graphphysics/dataset/test_xdmfdataset.py:31, 189-191).  The chunked + gzip + big-endian path is
exercised on the reference's airfoil sample when /root/reference is present (build container)."""
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
META = {"dt": 0.01, "features": {"velocity_x": {"type": "dynamic", "dtype": "float32"}, "velocity_y": {"type": "dynamic", "dtype": "float32"}}}


def test_hdf5_reader_on_the_reference_archive():
    from graphphysics_b200.io.hdf5 import H5File
    g = np.load(os.path.join(G, "cylinder_mesh.npz"))
    with H5File(os.path.join(G, "mock_xdmf", "mock.h5")) as f:
        assert sorted(f.keys(), key=lambda k: int(k[4:])) == [f"data{i}" for i in range(14)]
        assert f["data0"].shape == (1923, 3) and f["data0"].dtype == np.float32
        assert np.array_equal(f["data0"][()], g["points"])
        assert f["data1"].dtype == np.int64 and np.array_equal(f["data1"][()], g["triangles"])
        for t in range(6):
            assert np.array_equal(f[f"data{2 + 2 * t}"][()], g["velocity"][t, :, 0])
            assert np.array_equal(f[f"/data{3 + 2 * t}"][()], g["velocity"][t, :, 1])
        with pytest.raises(KeyError):
            f["nope"]


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/mock_airfoil/sample_000000005.h5"), reason="reference tree not present")
def test_hdf5_reader_chunked_gzip_big_endian():
    from graphphysics_b200.io.xdmf import TimeSeriesReader
    with TimeSeriesReader("/root/reference/tests/mock_airfoil/sample_000000005.xdmf") as r:
        points, cells = r.read_points_cells()
        t, pd, _ = r.read_data(0)
    assert points.shape == (27125, 3) and cells[0][0] == "triangle" and cells[0][1].shape == (52656, 3)
    assert cells[0][1].min() == 0 and cells[0][1].max() == 27124          # every node is used: the chunks were pasted at the right offsets
    assert set(pd) == {"Velocity_x", "Velocity_y", "Pressure", "Mach", "Distance", "Node_type"}
    assert all(v.shape == (27125,) and np.isfinite(v).all() for v in pd.values())
    assert set(np.unique(pd["Node_type"])) <= set(range(9))


def test_hdf5_write_read_round_trip(tmp_path):
    from graphphysics_b200.io.hdf5 import H5File, write_h5
    rng = np.random.default_rng(0)
    tree = {"a32": rng.random((7, 3)).astype(np.float32), "a64": rng.random((4, 5, 2)), "i32": rng.integers(-9, 9, (11,)).astype(np.int32),
            "i64": rng.integers(0, 1 << 40, (3, 2)), "u8": rng.integers(0, 255, (6,)).astype(np.uint8), "empty": np.zeros((0, 3), np.float32),
            "scalar": np.float32(3.5), "traj": {"velocity": rng.random((5, 9, 2)).astype(np.float32), "deep": {"cells": np.arange(12).reshape(4, 3)}}}
    for i in range(40):                                   # more members than a default symbol-table node holds
        tree[f"m{i:03d}"] = np.full((i % 5 + 1,), i, np.int16)
    p = str(tmp_path / "x.h5")
    write_h5(p, tree)
    with H5File(p) as f:
        assert set(f.keys()) == set(tree)
        for k, v in tree.items():
            if isinstance(v, dict):
                continue
            a = f[k][()]
            assert a.dtype == np.asarray(v).dtype and a.shape == np.asarray(v).shape and np.array_equal(a, v), k
        assert np.array_equal(f["traj"]["velocity"][()], tree["traj"]["velocity"])
        assert np.array_equal(f["traj/deep/cells"][()], tree["traj"]["deep"]["cells"])
        assert "traj" in f and "velocity" in f["traj"] and len(f["traj"]) == 2


def test_xdmf_time_series_reader_and_trajectory():
    from oracle import gp_oracle as O
    from graphphysics_b200.io.xdmf import TimeSeriesReader, XDMFTrajectory
    g = np.load(os.path.join(G, "cylinder_mesh.npz"))
    path = os.path.join(G, "mock_xdmf", "mock.xdmf")
    with TimeSeriesReader(path) as r:
        points, cells = r.read_points_cells()
        assert r.num_steps == 6 and cells[0][0] == "triangle"
        assert np.array_equal(points, g["points"]) and np.array_equal(cells[0][1], g["triangles"])
        for k in range(6):
            t, pd, cd = r.read_data(k)
            assert t == float(k) and not cd
            assert np.array_equal(pd["velocity_x"], g["velocity"][k, :, 0]) and np.array_equal(pd["velocity_y"], g["velocity"][k, :, 1])
    traj = XDMFTrajectory(path, META, targets=["velocity_x", "velocity_y"])
    assert len(traj) == 5 and traj.mesh_id == "mock"
    d = traj[2]
    # meshdata_to_graph (utils/torch_graph.py:137-221): x = [point data..., time], y = the next frame's targets, face (3, F)
    assert tuple(d.x.shape) == (1923, 3) and tuple(d.y.shape) == (1923, 2) and tuple(d.face.shape) == (3, 3612) and d.tetra is None
    assert torch.equal(d.x[:, :2], torch.from_numpy(g["velocity"][2])) and float(d.x[0, 2]) == 2.0
    assert torch.equal(d.y, torch.from_numpy(g["velocity"][3])) and d.pos.dtype == torch.float32
    assert O.face_to_edge(d.face.numpy().T, 1923).shape == (2, 11070)       # test_xdmfdataset.py:31
    with pytest.raises(IndexError):
        traj[5]


def test_xdmf_writer_round_trip_and_append(tmp_path):
    from graphphysics_b200.io.xdmf import TimeSeriesReader, append_frame_to_xdmf, meshes_to_xdmf
    from graphphysics_b200.synthetic import box_tet_mesh
    pos, tets = box_tet_mesh(4, 3, 3)
    rng = np.random.default_rng(1)
    frames = [{"velocity": rng.random((len(pos), 3)).astype(np.float32), "pressure": rng.random(len(pos)).astype(np.float32)} for _ in range(3)]
    base = str(tmp_path / "pred_7")
    meshes_to_xdmf(base, pos, [("tetra", tets)], frames, timestep=0.5)
    assert os.path.exists(base + ".xdmf") and os.path.exists(base + ".h5")
    extra = {"velocity": rng.random((len(pos), 3)).astype(np.float32), "pressure": rng.random(len(pos)).astype(np.float32)}
    append_frame_to_xdmf(base, extra, timestep=0.5)
    with TimeSeriesReader(base + ".xdmf") as r:
        p2, c2 = r.read_points_cells()
        assert np.array_equal(p2, pos) and c2[0][0] == "tetra" and np.array_equal(c2[0][1], tets) and r.num_steps == 4
        for k, fr in enumerate(frames + [extra]):
            t, pd, _ = r.read_data(k)
            assert t == 0.5 * k and np.array_equal(pd["velocity"], fr["velocity"]) and np.array_equal(pd["pressure"], fr["pressure"])


def test_h5_trajectory_file_layout(tmp_path):
    """The H5Dataset layout (utils/hierarchical.py:51-170): file[trajectory][feature], cast and reshaped by the meta JSON."""
    from graphphysics_b200.io.hdf5 import write_h5
    from graphphysics_b200.io.xdmf import H5Trajectories
    rng = np.random.default_rng(2)
    n, T = 30, 4
    cells = rng.integers(0, n, (50, 3)).astype(np.int32)
    meta = {"dt": 0.01, "features": {"cells": {"type": "static", "shape": [1, -1, 3], "dtype": "int32"},
                                     "mesh_pos": {"type": "static", "shape": [1, -1, 2], "dtype": "float32"},
                                     "node_type": {"type": "static", "shape": [1, -1, 1], "dtype": "int32"},
                                     "velocity": {"type": "dynamic", "shape": [T, -1, 2], "dtype": "float32"},
                                     "pressure": {"type": "dynamic", "shape": [T, -1, 1], "dtype": "float32"}}}
    trajs = {str(i): {"cells": cells[None], "mesh_pos": rng.random((1, n, 2)).astype(np.float32), "node_type": rng.integers(0, 7, (1, n, 1)).astype(np.int32),
                      "velocity": rng.random((T, n, 2)).astype(np.float32), "pressure": rng.random((T, n, 1)).astype(np.float32)} for i in range(3)}
    h5, mp = str(tmp_path / "train.h5"), str(tmp_path / "meta.json")
    write_h5(h5, trajs)
    json.dump(meta, open(mp, "w"))
    ds = H5Trajectories(h5, mp, targets=["velocity"])
    assert sorted(ds.keys) == ["0", "1", "2"] and len(ds) == 3
    d = ds.frame("1", 2)
    tr = trajs["1"]
    # point data order of get_frame_as_mesh: dynamic fields in file order, then node_type, then the time column
    exp = np.concatenate([tr["velocity"][2], tr["pressure"][2], tr["node_type"][0].astype(np.float32), np.full((n, 1), 2 * 0.01, np.float32)], 1)
    assert np.allclose(d.x.numpy(), exp) and np.array_equal(d.y.numpy(), tr["velocity"][3])
    assert np.array_equal(d.face.numpy(), cells.T) and np.array_equal(d.pos.numpy(), tr["mesh_pos"][0])
