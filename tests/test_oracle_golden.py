"""CPU: the oracle (oracle/gp_oracle.py, exact mode) against golden vectors produced by the
UNMODIFIED reference modules (oracle/make_golden.py, committed under tests/golden/), and against
the integer facts the reference's own tests pin."""
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O
from oracle.cpu_train import CpuTrainer
from tests.util import l2_rel

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(G, name))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    return z, sd


@pytest.mark.parametrize("name", ["epd_l2_h32.npz", "epd_l2_h64.npz"])
def test_epd_forward_and_gradients_match_reference(name):
    z, sd = _load(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.epd_forward(sd, torch.from_numpy(z["x"]), torch.from_numpy(z["edge_attr"]), torch.from_numpy(z["edge_index"]),
                        int(z["L"]))
    np.testing.assert_allclose(out.detach().numpy(), z["out"], rtol=1e-4, atol=1e-5)
    (out * torch.from_numpy(z["G"])).sum().backward()
    for k, p in sd.items():
        np.testing.assert_allclose(p.grad.numpy(), z["grad/" + k], rtol=2e-3, atol=2e-5, err_msg=k)


def test_graphnet_block_matches_reference():
    z, sd = _load("graphnet_block_h32.npz")
    sd = {"b." + k: v for k, v in sd.items()}
    ei = torch.from_numpy(z["edge_index"])
    ox, oe = O.graph_net_block(torch.from_numpy(z["x"]), torch.from_numpy(z["e"]), ei[0], ei[1], sd, "b")
    np.testing.assert_allclose(ox.numpy(), z["out_x"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(oe.numpy(), z["out_e"], rtol=1e-4, atol=1e-5)


def test_transformer_forward_and_gradients_match_reference():
    z, sd = _load("transformer_l2_h64.npz")
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.etd_forward(sd, torch.from_numpy(z["x"]), torch.from_numpy(z["edge_index"]), 2, 4)
    np.testing.assert_allclose(out.detach().numpy(), z["out"], rtol=1e-4, atol=1e-5)
    (out * torch.from_numpy(z["G"])).sum().backward()
    for k, p in sd.items():
        if p.grad is None:          # rope_inv_freq etc. are buffers, not in named_parameters
            continue
        np.testing.assert_allclose(p.grad.numpy(), z["grad/" + k], rtol=2e-3, atol=2e-5, err_msg=k)


def test_training_steps_match_reference():
    """Simulator (normalisers, one-hot, target delta) + L2Loss + clip + AdamW + cosine warm-up."""
    z = np.load(os.path.join(G, "train_steps.npz"))
    sd0 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0/")}
    index = dict(feature_index_start=0, feature_index_end=2, output_index_start=0, output_index_end=2, node_type_index=2)
    tr = CpuTrainer(sd0, 2, index, 2, 11, 3, lr=1e-3, num_steps=10, warmup=2)
    ei, ea = torch.from_numpy(z["edge_index"]), torch.from_numpy(z["edge_attr"])
    frames, ys = torch.from_numpy(z["frames"]), torch.from_numpy(z["ys"])
    for s in range(3):
        loss = tr.training_step(frames[s], ys[s], ea, ei)
        assert abs(loss - z["losses"][s]) <= 2e-5 * max(1.0, abs(z["losses"][s])), (s, loss, z["losses"][s])
        assert abs(tr.opt.param_groups[0]["lr"] - z["lrs"][s]) < 1e-12
    for k, p in tr.params.items():
        np.testing.assert_allclose(p.detach().numpy(), z["sd3/model." + k], rtol=2e-3, atol=2e-6, err_msg=k)
    np.testing.assert_allclose(tr.norms["node"].acc_sum.numpy(), z["sd3/_node_normalizer._acc_sum"], rtol=1e-5)
    np.testing.assert_allclose(tr.norms["output"].acc_sum_sq.numpy(), z["sd3/_output_normalizer._acc_sum_squared"], rtol=1e-5)
    with torch.no_grad():
        net, tgt, outp = tr.forward(frames[3], ys[3], ea, ei, training=False)
    np.testing.assert_allclose(net.numpy(), z["eval_net"], rtol=5e-3, atol=5e-5)
    np.testing.assert_allclose(tgt.numpy(), z["eval_target"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(outp.numpy(), z["eval_outputs"], rtol=5e-3, atol=5e-5)


def test_small_ops_match_reference():
    z = np.load(os.path.join(G, "small_ops.npz"))
    np.testing.assert_allclose(O.rms_norm(torch.from_numpy(z["rms_x"]), torch.ones(16)).numpy(), z["rms_out"], rtol=1e-6)
    nz = O.Normalizer(5)
    n1 = nz(torch.from_numpy(z["norm_d1"]))
    n2 = nz(torch.from_numpy(z["norm_d2"]))
    np.testing.assert_allclose(n1.numpy(), z["norm_n1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(n2.numpy(), z["norm_n2"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(nz.inverse(n2).numpy(), z["norm_inv"], rtol=1e-5, atol=1e-6)   # test_layers.py:92-100
    np.testing.assert_allclose([O.cosine_warmup_factor(e, 2, 10) for e in range(12)], z["sched"], rtol=1e-12)


def test_integer_goldens_of_the_reference_tests():
    """tests/graphphysics/dataset/test_xdmfdataset.py:31,46,189-191,247-249 of the reference."""
    c = np.load(os.path.join(G, "cylinder_mesh.npz"))
    assert c["points"].shape == (1923, 3) and c["triangles"].shape == (3612, 3)
    ei = O.face_to_edge(c["triangles"], 1923)
    assert ei.shape == (2, 11070)
    assert O.edge_features(c["points"], ei).shape == (11070, 4)
    assert O.khop_edges(ei, 1923, 2).shape == (2, 32638)
    assert (np.diff(ei[0] * 1923 + ei[1]) > 0).all()            # coalesced: sorted by (row, col), unique
    rev = np.stack([ei[1], ei[0]])
    assert set(map(tuple, rev.T)) == set(map(tuple, ei.T))      # symmetric
    a = np.load(os.path.join(G, "aneurysm_mesh.npz"))
    assert a["points"].shape == (22535, 3) and a["tets"].shape == (115275, 4)
    assert O.face_to_edge(O.tetra_to_faces(a["tets"]), 22535).shape == (2, 291144)


def test_l2_loss_masking():
    """tests/graphphysics/utils/test_loss.py:58-123 of the reference: only NORMAL/OUTFLOW rows count."""
    out, tgt = torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]), torch.zeros(3, 2)
    nt = torch.tensor([0.0, 6.0, 5.0])
    assert float(O.l2_loss(tgt, out, nt)) == pytest.approx((1 + 4 + 25 + 36) / 4)
    assert bool(O.boundary_mask(nt).tolist() == [False, True, False])


def test_rollout_matches_reference():
    """Autoregressive roll-out over the reference's mock cylinder trajectory (tests/golden/rollout.npz, made by
    oracle/make_golden_rollout.py with the reference's Simulator): the oracle's simulator_forward + rollout
    reproduce every predicted frame and both RMSE metrics (lightning_module.py:375-409, 446-486)."""
    z = np.load(os.path.join(G, "rollout.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    msd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    norms = {}
    for name, key, size in (("node", "_node_normalizer", 11), ("edge", "_edge_normalizer", 3), ("output", "_output_normalizer", 2)):
        norms[name] = O.Normalizer(size)
        norms[name].load(sd, key)
    index = dict(feature_index_start=0, feature_index_end=2, output_index_start=0, output_index_end=2, node_type_index=2)
    ei, ea = torch.from_numpy(z["edge_index"]), torch.from_numpy(z["edge_attr"])
    frames, ys = torch.from_numpy(z["frames"]), torch.from_numpy(z["ys"])

    def step(x_raw, y):
        model_fn = lambda nf, ef: O.epd_forward(msd, nf, ef, ei, 3, mode=None)
        return O.simulator_forward(model_fn, norms, x_raw, y, ea, index, training=False)[2]

    preds, r1, rall = O.rollout(step, list(frames), list(ys), 0, 2, 2)
    np.testing.assert_allclose(torch.stack(preds).numpy(), z["predictions"], rtol=1e-3, atol=1e-5)
    assert abs(r1 - float(z["val_1step_rmse"])) < 1e-5 and abs(rall - float(z["val_all_rollout_rmse"])) < 1e-5


def _bench_golden():
    """tests/golden/epd_l15_h128.npz (oracle/make_golden_bench.py): the weights are default_state_dict(seed=0),
    verified against the fixture's checksums before anything is compared."""
    from oracle.cpu_train import default_state_dict
    z = np.load(os.path.join(G, "epd_l15_h128.npz"))
    sd = default_state_dict(int(z["L"]), 11, 3, 2, int(z["H"]), seed=0)
    for k, v in sd.items():
        got = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        assert np.allclose(got, z["sdsum/" + k], rtol=1e-9, atol=1e-12), \
            f"default init of {k} is not the one the golden was made with (torch RNG / init changed): regenerate"
    return z, sd


def test_benchmark_config_l15_h128_matches_reference():
    """The oracle at the BENCHMARKED depth and width (BASELINE configs[1]: 15 layers, hidden 128) on one graph of the
    benchmark batch, against the reference's own modules: output, scalar, every gradient's norm, 16 full gradients."""
    torch.set_num_threads(8)
    z, sd = _bench_golden()
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.epd_forward(sd, torch.from_numpy(z["x"]), torch.from_numpy(z["edge_attr"]), torch.from_numpy(z["edge_index"]), 15)
    np.testing.assert_allclose(out.detach().numpy(), z["out"], rtol=2e-4, atol=2e-5)
    s = (out * torch.from_numpy(z["G"])).sum()
    assert abs(s.item() - float(z["scalar"])) < 1e-4
    s.backward()
    for k, p in sd.items():
        gn = float(z["gnorm/" + k])
        assert abs(p.grad.double().norm().item() - gn) <= 1e-3 * gn + 1e-7, k
        if "grad/" + k in z.files:
            ref = z["grad/" + k]
            assert np.linalg.norm(p.grad.numpy() - ref) <= 1e-3 * np.linalg.norm(ref) + 1e-7, k


def test_cylinder_json_verbatim_matches_reference():
    """training_config/cylinder.json verbatim on the reference's mock cylinder trajectory: three training steps and
    an eval step of the reference (tests/golden/cylinder_json_step.npz) reproduced by the oracle's CPU trainer."""
    import json
    cfg = json.load(open(os.path.join(G, "training_configs.json")))["cylinder"]
    m, index = cfg["model"], cfg["index"]
    z = np.load(os.path.join(G, "cylinder_json_step.npz"))
    sd0 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0/")}
    tr = CpuTrainer(sd0, m["message_passing_num"], index, m["output_size"], m["node_input_size"] + 9, m["edge_input_size"],
                    lr=1e-3, num_steps=10, warmup=2)
    ei, ea = torch.from_numpy(z["edge_index"]), torch.from_numpy(z["edge_attr"])
    frames, ys = torch.from_numpy(z["frames"]), torch.from_numpy(z["ys"])
    for s in range(3):
        loss = tr.training_step(frames[s], ys[s], ea, ei)
        assert abs(loss - z["losses"][s]) <= 2e-5 * max(1.0, abs(z["losses"][s])), (s, loss, z["losses"][s])
    with torch.no_grad():
        net, tgt, outp = tr.forward(frames[3], ys[3], ea, ei, training=False)
    np.testing.assert_allclose(net.numpy(), z["eval_net"], rtol=5e-3, atol=5e-5)
    np.testing.assert_allclose(outp.numpy(), z["eval_outputs"], rtol=5e-3, atol=5e-5)


def test_oracle_world_edges_against_brute_force():
    """oracle.world_edges calls scipy's cKDTree like the reference (preprocessing.py:112-117); pin it against a brute-force
    fp64 all-pairs evaluation of the same definition (distance <= radius, OBSTACLE-NORMAL ends, undirected, coalesced)."""
    from oracle import gp_oracle as O
    rng = np.random.default_rng(0)
    n = 600
    pos = (rng.random((n, 3)) * np.array([0.5, 0.5, 0.1])).astype(np.float32)
    t = rng.integers(0, 3, n)
    mesh = np.stack([rng.integers(0, n, 200), rng.integers(0, n, 200)])
    got = O.world_edges(mesh, pos, t, n, 0.03)
    p64 = pos.astype(np.float64)
    d2 = ((p64[:, None, :] - p64[None, :, :]) ** 2).sum(-1)
    i, j = np.nonzero((d2 <= 0.03 ** 2) & (t[:, None] == O.OBSTACLE) & (t[None, :] == O.NORMAL))
    row = np.concatenate([i, j, mesh[0], mesh[1]])
    col = np.concatenate([j, i, mesh[1], mesh[0]])
    key = np.unique(row.astype(np.int64) * n + col)
    assert np.array_equal(got, np.stack([key // n, key % n]))


EPD_VARIANTS = {"epd_silu": dict(act="silu"), "epd_gated_mlp": dict(gated_mlp=True), "epd_gated_mlp_silu": dict(act="silu", gated_mlp=True),
                "epd_gate": dict(gate=True), "epd_rope": dict(rope_axes=2),
                "epd_all": dict(act="silu", gated_mlp=True, gate=True, rope_axes=2), "epd_temporal": dict(temporal=True)}
ETD_VARIANTS = {"etd_gated_attention": dict(gated_attention=True), "etd_rope": dict(rope=True), "etd_silu": dict(act="silu"),
                "etd_shared_qkv": dict(), "etd_temporal": dict(temporal=True)}


def _variant_case(z, name):
    sd = {k[len(name) + 4:]: torch.from_numpy(z[k]).double().requires_grad_(True) for k in z.files if k.startswith(name + "/sd/")}
    grads = {k[len(name) + 6:]: z[k] for k in z.files if k.startswith(name + "/grad/")}
    return sd, grads


def test_oracle_variant_flags_against_reference_golden():
    """SiLU / gated MLP / aggregation gate / relative RoPE (EncodeProcessDecode) and gated attention / RoPE / SiLU / shared
    q-k-v weights (EncodeTransformDecode), and the temporal block (TemporalAttention) of both: the oracle reproduces the UNMODIFIED reference's outputs (1e-5) and parameter
    gradients (1e-4; tests/golden/variants.npz, oracle/make_golden_variants.py)."""
    from oracle import gp_oracle as O
    z = np.load(os.path.join(G, "variants.npz"))
    ei, ea, pos = torch.from_numpy(z["edge_index"]), torch.from_numpy(z["edge_attr"]).double(), torch.from_numpy(z["pos"]).double()
    for name, kw in EPD_VARIANTS.items():
        sd, grads = _variant_case(z, name)
        out = O.epd_forward_variant(sd, torch.from_numpy(z["x_epd"]).double(), ea, ei, 2, pos=pos, phi=torch.from_numpy(z["phi"]).double(), **kw)
        assert l2_rel(out, torch.from_numpy(z[name + "/out"])) < 1e-5, name
        (out * torch.from_numpy(z["G_epd"]).double()).sum().backward()
        biggest = max(float(np.linalg.norm(g)) for g in grads.values())
        for k, g in grads.items():
            if float(np.linalg.norm(g)) < 1e-7 * biggest:                 # temporal_block.k_proj.bias: analytically zero
                continue
            assert l2_rel(sd[k].grad, torch.from_numpy(g)) < 2e-4, (name, k)
    for name, kw in ETD_VARIANTS.items():
        sd, grads = _variant_case(z, name)
        shared = name == "etd_shared_qkv"
        out = O.etd_forward_variant(sd, torch.from_numpy(z["x_etd"]).double(), ei, 2, 4, pos=pos, **kw)
        assert l2_rel(out, torch.from_numpy(z[name + "/out"])) < 1e-5, name
        (out * torch.from_numpy(z["G_etd"]).double()).sum().backward()
        biggest = max(float(np.linalg.norm(g)) for g in grads.values())
        for k, g in grads.items():
            if float(np.linalg.norm(g)) < 1e-7 * biggest:                 # k_proj.bias without RoPE: analytically zero
                continue
            got = sd[k].grad
            if shared and k.endswith("attention.q_proj.weight"):       # one Parameter behind q / k / v: its gradient is the sum
                got = got + sd[k.replace("q_proj", "k_proj")].grad + sd[k.replace("q_proj", "v_proj")].grad
            assert l2_rel(got, torch.from_numpy(g)) < 2e-4, (name, k)


def test_oracle_attention_values_against_reference_golden():
    """return_attention=True (layers.py:493-522, 680-683, 795-801): the oracle's sparse-softmax values equal attn.val of the
    UNMODIFIED reference's first Transformer block, in the caller's edge order, with and without RoPE
    (tests/golden/variants.npz, oracle/make_golden_variants.py)."""
    from oracle import gp_oracle as O
    z = np.load(os.path.join(G, "variants.npz"))
    ei, pos = torch.from_numpy(z["edge_index"]), torch.from_numpy(z["pos"]).double()
    for name, rope in (("etd_gated_attention", False), ("etd_rope", True)):
        sd = {k[len(name) + 4:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith(name + "/sd/")}
        h0 = torch.from_numpy(z[name + "/h0"]).double()
        N, H, heads = h0.shape[0], h0.shape[1], 4
        d = H // heads
        p = "processor_list.0"
        n1 = O.rms_norm(h0, sd[f"{p}.norm1.scale"])
        q = O.linear(n1, sd[f"{p}.attention.q_proj.weight"], sd[f"{p}.attention.q_proj.bias"], None).reshape(N, d, heads)
        k = O.linear(n1, sd[f"{p}.attention.k_proj.weight"], sd[f"{p}.attention.k_proj.bias"], None).reshape(N, d, heads)
        if rope:
            q, k = O.rope_nodes(q, k, pos, sd[f"{p}.attention.rope_inv_freq"])
        got = O.attention_values(q, k, ei[0], ei[1], N)
        ref = torch.from_numpy(z[name + "/attn0"]).double()
        assert got.shape == ref.shape and float((got - ref).abs().max()) < 2e-6, name
