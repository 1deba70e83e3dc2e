"""CPU: host-side logic of graphphysics_b200 (no compute call reaches the GPU library)."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from graphphysics_b200 import _lib
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "gp_b200.h")).read()
    names = sorted(set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libgp_b200.so does not export {n}"
    assert lib.gp_version() >= 100


def test_product_refuses_to_run_without_cuda():
    """No CPU / PyTorch fallback: building the engine on CPU tensors raises."""
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = EncodeProcessDecode(1, 11, 3, 2, hidden_size=32)
    g = Data(x=torch.zeros(4, 11), edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(g)


def test_state_dict_keys_match_reference_layout():
    """same keys and shapes as the reference's EncodeProcessDecode."""
    from graphphysics_b200.models.processors import EncodeProcessDecode
    z = np.load(os.path.join(ROOT, "tests", "golden", "epd_l2_h32.npz"))
    ref = {k[3:]: z[k].shape for k in z.files if k.startswith("sd/")}
    mine = {k: tuple(v.shape) for k, v in EncodeProcessDecode(2, 11, 3, 2, hidden_size=32).state_dict().items()}
    assert mine == ref


def test_graph_csr_bit_exact_against_oracle():
    from graphphysics_b200.graph import GraphCSR
    rng = np.random.default_rng(0)
    n, e = 57, 400
    ei = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)])
    ei[1, :30] = 7                                    # a long receiver segment
    ref = O.csr_by_receiver(ei, n)
    g = GraphCSR(torch.from_numpy(ei), n)
    assert np.array_equal(g.perm_dst.numpy(), ref["perm_dst"])
    assert np.array_equal(g.rowptr_dst.numpy(), ref["rowptr_dst"])
    assert np.array_equal(g.rowptr_src.numpy(), ref["rowptr_src"])
    assert np.array_equal(g.dst.numpy(), ei[1][ref["perm_dst"]])
    src_sorted = ei[0][ref["perm_dst"]]
    assert np.array_equal(g.perm_src.numpy(), np.argsort(src_sorted, kind="stable"))
    # empty graph
    g0 = GraphCSR(torch.zeros((2, 0), dtype=torch.long), 5)
    assert g0.num_edges == 0 and g0.rowptr_dst.tolist() == [0] * 6


def test_synthetic_mesh_edges_follow_face_to_edge():
    from graphphysics_b200 import synthetic as S
    c = np.load(os.path.join(ROOT, "tests", "golden", "cylinder_mesh.npz"))
    ei = S.mesh_edges(c["triangles"], 1923)
    assert ei.shape == (2, 11070) and np.array_equal(ei, O.face_to_edge(c["triangles"], 1923))
    assert np.allclose(S.mesh_edge_attr(c["points"], ei), O.edge_features(c["points"], ei))
    a = np.load(os.path.join(ROOT, "tests", "golden", "aneurysm_mesh.npz"))
    assert S.mesh_edges(S.faces_of_cells(a["tets"]), 22535).shape == (2, 291144)
    b = S.cylinder_flow_batch(2, nx=20, ny=10)
    n = b.x.shape[0]
    assert b.edge_index.max() < n and b.edge_attr.shape == (b.edge_index.shape[1], 3) and b.y.shape == (n, 2)


def test_scheduler_and_normalizer_modules_match_oracle():
    from graphphysics_b200.models.layers import Normalizer
    from graphphysics_b200.utils.scheduler import CosineWarmupScheduler, lr_factor
    for e in range(-1, 40):
        assert lr_factor(e, 5, 30) == pytest.approx(O.cosine_warmup_factor(e, 5, 30), rel=1e-12)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = CosineWarmupScheduler(opt, warmup=5, max_iters=30)
    lrs = []
    for _ in range(8):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    assert lrs == pytest.approx([O.cosine_warmup_factor(e, 5, 30) for e in range(8)])
    nz, ref = Normalizer(4, device="cpu"), O.Normalizer(4)
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        d = torch.randn(6, 4, generator=g)
        assert torch.allclose(nz(d), ref(d), atol=1e-6)
    assert torch.allclose(nz.inverse(nz(d, accumulate=False)), d, atol=1e-5)      # reference test_layers.py:92-100
    nz2 = Normalizer(4, device="cpu")
    nz2.load_state_dict(nz.state_dict())
    assert torch.allclose(nz2(d, accumulate=False), nz(d, accumulate=False))


def test_bf16_ulp_flip_amplification():
    """Why model-level GPU parity is looser than kernel-level parity: tipping ~1 % of the bf16
    roundings of the encoder outputs by one ulp (what fp32 summation order does) moves the output of
    a 3-layer default-init model by several 1e-3 -- the RMSNorm of each MLP rescales its small
    pre-norm output to unit RMS.  Measured on the oracle alone (kernel arithmetic mode)."""
    from oracle.cpu_train import default_state_dict
    torch.manual_seed(0)
    pos, tris = O.grid_tri_mesh(16, 10, jitter=0.3, seed=0)
    ei = torch.from_numpy(O.face_to_edge(tris, len(pos)))
    ea = torch.from_numpy(O.edge_features(pos, ei.numpy())).double()
    sd = default_state_dict(3, 11, 3, 2, 64, seed=1, dtype=torch.float64)
    x = torch.randn(len(pos), 11).double()
    x0 = O.rnd(O.mlp(x, sd, "nodes_encoder", mode="bf16"), "bf16")
    e0 = O.rnd(O.mlp(ea, sd, "edges_encoder", mode="bf16"), "bf16")

    def run(xx, ee):
        for i in range(3):
            xx, ee = O.graph_net_block(xx, ee, ei[0], ei[1], sd, f"processor_list.{i}", "bf16")
        return O.mlp(xx, sd, "decode_module", layer_norm=False, mode="bf16")

    g = torch.Generator().manual_seed(1)

    def flip(t):
        m = torch.rand(t.shape, generator=g) < 0.01
        ulp = t.abs().clamp_min(1e-30).log2().floor().exp2() * 2 ** -7
        return torch.where(m, t + ulp * torch.where(torch.rand(t.shape, generator=g) < 0.5, 1.0, -1.0), t)

    base, pert = run(x0, e0), run(flip(x0), flip(e0))
    injected = float((flip(x0) - x0).norm() / x0.norm())
    moved = float((pert - base).norm() / base.norm())
    assert injected < 1.5e-3
    assert 5e-4 < moved < 3e-2, moved


def test_spec_drift_vs_exact_arithmetic_is_bounded():
    """bf16 kernel arithmetic vs the reference's exact arithmetic on the same weights: the drift the
    SURVEY (Appendix B) measured for bf16 operands, here including bf16 storage of the latents."""
    from oracle.cpu_train import default_state_dict
    torch.manual_seed(0)
    pos, tris = O.grid_tri_mesh(16, 10, jitter=0.3, seed=0)
    ei = torch.from_numpy(O.face_to_edge(tris, len(pos)))
    ea = torch.from_numpy(O.edge_features(pos, ei.numpy())).double()
    x = torch.randn(len(pos), 11).double()
    sd = default_state_dict(5, 11, 3, 2, 32, seed=0, dtype=torch.float64)
    exact = O.epd_forward(sd, x, ea, ei, 5, mode=None)
    spec = O.epd_forward(sd, x, ea, ei, 5, mode="bf16")
    assert float((spec - exact).norm() / exact.norm()) < 2e-2


def test_shipped_training_configs_parse_verbatim():
    """get_model / get_simulator read the `model` / `index` / `training` sections of the reference's
    training_config/*.json unchanged (parse_parameters.py:81-190); plate.json is a transformer."""
    import json
    from graphphysics_b200.models.processors import EncodeProcessDecode, EncodeTransformDecode
    from graphphysics_b200.training.parse_parameters import get_model, get_simulator
    cfgs = json.load(open(os.path.join(ROOT, "tests", "golden", "training_configs.json")))
    m = get_model(cfgs["cylinder"])
    assert isinstance(m, EncodeProcessDecode) and len(m.processor_list) == 5 and m.hidden_size == 32
    assert m.nodes_encoder[0].in_features == 2 + 9 and m.edges_encoder[0].in_features == 3
    for name, nin in (("plate", 6), ("coarse-aneurysm", 14)):
        t = get_model(cfgs[name])
        assert isinstance(t, EncodeTransformDecode) and len(t.processor_list) == 10
        assert t.nodes_encoder[0].in_features == nin + 9 and t.decode_module[6].out_features == 3
        sim = get_simulator(cfgs[name], t, torch.device("cpu"))
        assert sim._edge_normalizer is None and sim.node_input_size == nin + 9
    with pytest.raises(ValueError):
        get_model({"model": {"type": "nope", "node_input_size": 1}})
    # variant flags build the reference's modules and route to the general path
    from graphphysics_b200.models import layers as L
    var = json.loads(json.dumps(cfgs["cylinder"]))
    var["model"].update(use_silu_activation=True, use_gated_mlp=True, use_gated_attention=True, use_rope_embeddings=True,
                        rope_pos_dimension=2)
    try:
        v = get_model(var)
        assert v.variant and isinstance(v.nodes_encoder[1], torch.nn.SiLU)
        keys = set(v.state_dict().keys())
        assert {"processor_list.0.gate_proj.weight", "processor_list.0.gate_pos", "processor_list.0.edge_block.1.linear1.weight",
                "processor_list.0.edge_block.0.scale", "processor_list.0.node_block.2.bias"} <= keys
        assert v.processor_list[0].edge_block[0].scale.numel() == 3 * 32 and v.processor_list[0]._pair_count == 32 // 4
    finally:
        L.set_use_silu_activation(False)
    assert not get_model(cfgs["cylinder"]).variant
    tv = json.loads(json.dumps(cfgs["coarse-aneurysm"]))
    tv["model"].update(use_gated_attention=True, use_rope_embeddings=True)
    t = get_model(tv)
    assert t.processor_list[0].attention.gate_proj is not None and t.processor_list[0].attention.m == 16 // 6
    assert "processor_list.0.attention.rope_inv_freq" in t.state_dict()
    for base in ("cylinder", "coarse-aneurysm"):                        # training.use_temporal_block (parse_parameters.py:103)
        tb = json.loads(json.dumps(cfgs[base]))
        tb["training"] = {"use_temporal_block": True}
        m = get_model(tb)
        assert {"temporal_block.q_proj.weight", "temporal_block.out_proj.bias", "temporal_block.gate.0.weight",
                "temporal_block.gate.2.bias", "temporal_block.mixer.0.weight", "temporal_block.mixer.2.bias"} <= set(m.state_dict().keys())
        assert m.temporal_block.gate[0].weight.shape == (m.hidden_size, 2 * m.hidden_size)


def test_kuhn_box_graph_equals_tetra_mesh_edges():
    """The direct generator of the structured tetrahedral mesh graph (used for the 1M-node runs) gives
    exactly the edge list of FaceToEdge on the 6-tets-per-cell mesh (torch_graph.py:194-210 semantics)."""
    import numpy as np
    from graphphysics_b200.synthetic import box_tet_mesh, faces_of_cells, kuhn_box_graph, mesh_edges
    for shape in ((6, 5, 4), (3, 3, 3), (2, 7, 3)):
        pos, ei = kuhn_box_graph(*shape)
        p2, tets = box_tet_mesh(*shape)
        assert np.array_equal(pos, p2)
        assert np.array_equal(ei, mesh_edges(faces_of_cells(tets), len(p2)))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU and prints
    one JSON line with the contract keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from oracle import build_ref
    build_ref.main()                   # stages the reference modules when /root/reference is present (build container)
    staged = os.path.exists(os.path.join(root, "oracle", "_ref", "MANIFEST.json"))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-graphs", "2"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()[-1]
    line = json.loads(out)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")      # the reference's own modules when staged
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_get_preprocessing_and_loss_from_the_json_configs():
    """parse_parameters.get_preprocessing / get_loss / get_gradient_method (parse_parameters.py:24-78, 300-341) on the
    reference's own training configs: the transform pipeline in the reference's order (noise second, world edges for the
    plate), L2Loss for the configurations of the path."""
    import json
    from graphphysics_b200.training.parse_parameters import get_gradient_method, get_loss, get_preprocessing
    from graphphysics_b200.utils.loss import L2Loss
    cfgs = json.load(open(os.path.join(ROOT, "tests", "golden", "training_configs.json")))
    # the `transformations` sections of training_config/cylinder.json and plate.json, verbatim
    cfgs["cylinder"]["transformations"] = {"preprocessing": {"noise": 0.02, "noise_index_start": [0], "noise_index_end": [2], "masking": 0},
                                           "world_pos_parameters": {"use": False, "world_pos_index_start": 0, "world_pos_index_end": 3}}
    cfgs["plate"]["transformations"] = {"preprocessing": {"noise": 0.003, "noise_index_start": [0], "noise_index_end": [3], "masking": 0},
                                        "world_pos_parameters": {"use": True, "world_pos_index_start": 0, "world_pos_index_end": 3}}
    names = lambda pre: [getattr(t, "func", t).__name__ for t in pre.transforms]
    cyl = get_preprocessing(cfgs["cylinder"], torch.device("cpu"))
    assert names(cyl) == ["face_to_edge", "add_noise", "_apply"]
    assert cyl.transforms[1].keywords["noise_scale"] == cfgs["cylinder"]["transformations"]["preprocessing"]["noise"]
    assert names(get_preprocessing(cfgs["cylinder"], torch.device("cpu"), remove_noise=True)) == ["face_to_edge", "_apply"]
    assert names(get_preprocessing(cfgs["cylinder"], torch.device("cpu"), use_edge_feature=False, remove_noise=True)) == ["face_to_edge"]
    plate = get_preprocessing(cfgs["plate"], torch.device("cpu"))
    assert names(plate) == ["add_obstacles_next_pos", "add_noise", "face_to_edge", "add_world_edges", "_apply"]
    loss, name = get_loss(cfgs["cylinder"])
    assert isinstance(loss, L2Loss) and name == "L2LOSS" and get_gradient_method(cfgs["cylinder"]) is None
    with pytest.raises(NotImplementedError):
        get_loss({"loss": {"type": ["l2loss", "divergenceloss"], "weights": [1.0, 0.1]}})


def test_collate_builds_the_union_graph_like_pyg():
    """graph.collate: node / edge tensors concatenated, edge_index and face shifted by the node offsets, batch and ptr
    vectors -- the block-diagonal union graph PyG's DataLoader hands to the reference's training step."""
    from graphphysics_b200.graph import Data, collate
    g1 = Data(x=torch.arange(6.).view(3, 2), y=torch.ones(3, 2), pos=torch.zeros(3, 2), edge_index=torch.tensor([[0, 1, 2], [1, 2, 0]]),
              edge_attr=torch.arange(9.).view(3, 3), face=torch.tensor([[0], [1], [2]]), traj_index=torch.tensor(4), name="a")
    g2 = Data(x=torch.arange(8.).view(4, 2) + 10, y=torch.zeros(4, 2), pos=torch.ones(4, 2),
              edge_index=torch.tensor([[0, 3], [3, 0]]), edge_attr=torch.ones(2, 3), face=torch.tensor([[0, 1], [1, 2], [3, 3]]),
              traj_index=torch.tensor(7), name="b")
    u = collate([g1, g2])
    assert u.x.shape == (7, 2) and u.y.shape == (7, 2) and u.pos.shape == (7, 2) and u.edge_attr.shape == (5, 3)
    assert torch.equal(u.edge_index, torch.tensor([[0, 1, 2, 3, 6], [1, 2, 0, 6, 3]]))
    assert torch.equal(u.face, torch.tensor([[0, 3, 4], [1, 4, 5], [2, 6, 6]]))
    assert torch.equal(u.batch, torch.tensor([0, 0, 0, 1, 1, 1, 1])) and torch.equal(u.ptr, torch.tensor([0, 3, 7]))
    assert torch.equal(u.traj_index, torch.tensor([4, 7])) and u.name == ["a", "b"]
    assert torch.equal(u.x[3:], g2.x) and u.num_nodes == 7
    one = collate([g1])
    assert torch.equal(one.edge_index, g1.edge_index) and torch.equal(one.batch, torch.zeros(3, dtype=torch.long))


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ (what `python bench.py [--gpus N]` printed on the B200 box) carry every key of
    the benchmark contract, with consistent values."""
    import json
    for n in (1, 2, 4, 8):
        path = os.path.join(ROOT, "profiles", f"r02_bench_{n}gpu.json")
        d = json.loads(open(path).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert "workload" in d["config"] and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        # value = directed edges x MP layers x ranks / step time
        edges = d["config"]["directed_edges_per_gpu"] * d["config"]["mp_layers"] * n
        assert abs(d["value"] - edges / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
        if n == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "reference"
        else:
            assert d["partition"]["scaling"] == "strong" and d["partition"]["n_gpus"] == n
