"""Model-level GPU parity: EncodeProcessDecode / GraphNetBlock on the CUDA kernels vs the CPU
oracle (kernel-arithmetic mode, fp64 accumulate) on identical weights and inputs.

Three levels of evidence:
  * teacher-forced, per layer: every GraphNetBlock is run on the oracle's own layer inputs and
    compared with the oracle's layer outputs -- l2 <= 1e-3 (what one layer adds: fp32 summation
    order plus the ~1 % of bf16 roundings it tips by one ulp);
  * free-running forward: l2 <= 3e-2.  A one-ulp tip in layer l is amplified by the RMSNorm of the
    following MLPs (their pre-norm outputs have RMS << 1 at default init), about 5x through the
    next layer; tests/test_oracle_cpu.py::test_bf16_ulp_flip_amplification measures that on the
    oracle alone.  The drift between the bf16 kernel arithmetic and exact arithmetic is printed
    beside it -- it is the same order;
  * free-running gradients: l2 <= 0.3, a guard against missing terms or wrong signs only (the same
    amplification applies twice); the tight gradient parity is in tests/test_mlp_bwd_gpu.py and in
    test_graphnet_block_standalone_api below."""
import numpy as np
import pytest
import torch

from tests.util import check_close, l2_rel

pytestmark = pytest.mark.gpu


def _mesh_graph(nx, ny, seed=0):
    from oracle import gp_oracle as O
    pos, tris = O.grid_tri_mesh(nx, ny, jitter=0.3, seed=seed, hole=(0.4, 0.2, 0.08))
    ei = O.face_to_edge(tris, pos.shape[0])
    ea = O.edge_features(pos, ei)
    return pos, torch.from_numpy(ei), torch.from_numpy(ea)


@pytest.mark.parametrize("hidden,layers", [(32, 2), (64, 2), (128, 3)])
def test_epd_forward_backward_matches_oracle(hidden, layers):
    from oracle import gp_oracle as O
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    dev = torch.device("cuda:0")
    torch.manual_seed(hidden)
    pos, ei, ea = _mesh_graph(24, 14)
    N, E = pos.shape[0], ei.shape[1]
    x = torch.randn(N, 11)
    model = EncodeProcessDecode(layers, 11, 3, 2, hidden_size=hidden)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    G = torch.randn(N, 2)

    # oracle, kernel arithmetic
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    ref = O.epd_forward(sd64, x.double(), ea.double(), ei, layers, mode="bf16")
    (ref * G.double()).sum().backward()
    # oracle, reference arithmetic (for the drift report only)
    exact = O.epd_forward({k: v.double() for k, v in sd.items()}, x.double(), ea.double(), ei, layers, mode=None)

    graph = Data(x=x.to(dev), edge_index=ei.to(dev), edge_attr=ea.to(dev))
    out = model(graph)
    (out * G.to(dev)).sum().backward()
    torch.cuda.synchronize()
    grads = model.engine.grads_by_name()

    rep, ok = [], True
    ok &= check_close(out, ref, "forward vs kernel-spec", 3e-2, 1e-1, rep)
    rep.append(f"    drift of bf16 kernel arithmetic vs exact fp64 reference arithmetic: l2_rel={l2_rel(out, exact):.2e}")
    for name, gref in sd64.items():
        ok &= check_close(grads[name], gref.grad, name, 3e-1, 1.0, rep)

    # teacher-forced per-layer forward
    from graphphysics_b200 import ops
    from graphphysics_b200.graph import get_csr
    eng = model.engine
    g = get_csr(graph.edge_index, N)
    perm = g.perm_dst64.cpu()
    sdd = {k: v.detach() for k, v in sd64.items()}
    with torch.no_grad():
        xo = O.rnd(O.mlp(x.double(), sdd, "nodes_encoder", mode="bf16"), "bf16")
        eo = O.rnd(O.mlp(ea.double(), sdd, "edges_encoder", mode="bf16"), "bf16")
        bnd = torch.empty(ops.seg_bnd_size(E, hidden), dtype=torch.float32, device=dev)
        for l in range(layers):
            xn, en = O.graph_net_block(xo, eo, ei[0], ei[1], sdd, f"processor_list.{l}", mode="bf16")
            xk, ek, _ = eng.run_block(l, xo.to(dev).to(torch.bfloat16), eo[perm].to(dev).to(torch.bfloat16).contiguous(),
                                      g, bnd, False)
            ok &= check_close(xk.float(), xn, f"layer {l} x (teacher-forced)", 1e-3, 1e-2, rep)
            ok &= check_close(ek.float(), en[perm], f"layer {l} e (teacher-forced)", 1e-3, 1e-2, rep)
            xo, eo = xn, en
    print("\n".join(rep))
    assert ok, "\n".join(r for r in rep if r.startswith("BAD"))
    # the nn.Parameters are views of the engine's flat buffer: state_dict keys/shapes unchanged
    assert set(model.state_dict().keys()) == set(sd.keys())
    # eval / no-grad path gives the same numbers
    with torch.no_grad():
        out2 = model(graph)
    assert torch.equal(out2, out.detach())


def test_graphnet_block_standalone_api():
    """GraphNetBlock.forward(x, edge_index, edge_attr) -> (x, edge_attr) in the caller's edge order,
    with gradients for both latent inputs (layers.py:989-1042)."""
    from oracle import gp_oracle as O
    from graphphysics_b200.models.layers import GraphNetBlock
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    H = 64
    pos, ei, _ = _mesh_graph(16, 10, seed=2)
    N, E = pos.shape[0], ei.shape[1]
    ei = ei[:, torch.randperm(E)]                      # arbitrary (not receiver-sorted) edge order
    blk = GraphNetBlock(H)
    sd = {"processor_list.0." + k: v.detach().clone().double().requires_grad_(True) for k, v in blk.state_dict().items()}
    blk = blk.to(dev)
    x = torch.randn(N, H).to(torch.bfloat16).float()
    e = torch.randn(E, H).to(torch.bfloat16).float()
    x64, e64 = x.double().requires_grad_(True), e.double().requires_grad_(True)
    rx, re = O.graph_net_block(x64, e64, ei[0], ei[1], sd, "processor_list.0", mode="bf16")
    Gx, Ge = torch.randn(N, H), torch.randn(E, H).to(torch.bfloat16).float()
    ((rx * Gx.double()).sum() + (re * Ge.double()).sum()).backward()

    xd, ed = x.to(dev).requires_grad_(True), e.to(dev).requires_grad_(True)
    ox, oe = blk(xd, ei.to(dev), ed)
    ((ox * Gx.to(dev)).sum() + (oe * Ge.to(dev)).sum()).backward()
    torch.cuda.synchronize()
    rep, ok = [], True
    ok &= check_close(ox, rx, "x out", 3e-3, 1e-2, rep)
    ok &= check_close(oe, re, "e out", 3e-3, 1e-2, rep)
    ok &= check_close(xd.grad, x64.grad, "dx", 5e-3, 5e-2, rep)
    ok &= check_close(ed.grad, e64.grad, "de", 5e-3, 5e-2, rep)
    print("\n".join(rep))
    assert ok, "\n".join(rep)


def test_saved_activations_equal_recompute(monkeypatch):
    """The backward recomputes h1 / h3 from the layer inputs and the gathered pre-activations (default)
    or reads them back from the forward (GP_B200_SAVE_ALL=1): the recomputed tiles repeat the
    forward's instruction sequence, so both modes must give the same gradients (to fp32 rounding of
    the differently ordered partial sums: 1e-5)."""
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    pos, ei, ea = _mesh_graph(30, 17)
    N = pos.shape[0]
    x, G = torch.randn(N, 11), torch.randn(N, 2)
    grads = []
    sd = None
    for save_all in ("1", "0"):
        monkeypatch.setenv("GP_B200_SAVE_ALL", save_all)
        model = EncodeProcessDecode(3, 11, 3, 2, hidden_size=128)
        if sd is None:
            sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        model.load_state_dict(sd)
        model = model.to(dev)
        assert model.engine.save_all == (save_all == "1")
        out = model(Data(x=x.to(dev), edge_index=ei.to(dev), edge_attr=ea.to(dev)))
        (out * G.to(dev)).sum().backward()
        torch.cuda.synchronize()
        grads.append({k: v.detach().float().cpu() for k, v in model.engine.grads_by_name().items()})
    for k in grads[0]:
        assert l2_rel(grads[0][k], grads[1][k]) < 1e-5, k


@pytest.mark.parametrize("n_nodes,n_edges", [(5, 8), (40, 1), (130, 129), (300, 384), (200, 641), (7, 0)])
@pytest.mark.parametrize("hidden", [128, 32])
def test_epd_edge_case_graphs(n_nodes, n_edges, hidden):
    """Graphs smaller than one 128-row tile, exactly on tile boundaries, with an odd number of edge tiles (the CTA-pair
    forward kernel then processes an all-padding tile), with isolated nodes, with a single edge and with NO edges: the
    fused model equals the oracle in kernel mode (the free-running forward bound of this file) and the gradients are
    finite; nodes without in-edges aggregate zeros."""
    from oracle import gp_oracle as O
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(n_nodes * 1000 + n_edges)
    used = n_nodes - 3 if n_nodes >= 8 else n_nodes             # the last three nodes are isolated (not on the 5-node graph)
    key = rng.choice(used * used, size=n_edges, replace=False) if n_edges else np.zeros(0, np.int64)
    ei = torch.from_numpy(np.stack([key // used, key % used]).astype(np.int64))
    torch.manual_seed(hidden + n_edges)
    x, ea, G = torch.randn(n_nodes, 11), torch.randn(n_edges, 3), torch.randn(n_nodes, 2)
    model = EncodeProcessDecode(2, 11, 3, 2, hidden_size=hidden)
    sd64 = {k: v.detach().clone().double() for k, v in model.state_dict().items()}
    ref = O.epd_forward(sd64, x.double(), ea.double(), ei, 2, mode="bf16")
    model = model.to(dev)
    out = model(Data(x=x.to(dev), edge_index=ei.to(dev), edge_attr=ea.to(dev)))
    assert tuple(out.shape) == (n_nodes, 2)
    assert l2_rel(out, ref) < 3e-2, l2_rel(out, ref)
    (out * G.to(dev)).sum().backward()
    grads = model.engine.grads_by_name()
    assert all(torch.isfinite(g).all() for g in grads.values())
    if n_edges == 0:                                            # no message ever reaches the edge MLPs
        assert float(grads["processor_list.0.edge_block.6.weight"].abs().max()) == 0.0


def test_side_stream_scheduling_does_not_change_results(monkeypatch):
    """The engine runs the per-layer gradient reduction, the receiver-side fix-up and the node encoder on side streams
    (GP_B200_SIDE_REDUCE, default on).  Same kernels, same order of every sum: output and gradients are bit-identical to the
    single-stream schedule (the captured-graph form of the same schedule is what tests/test_trainer_gpu.py replays)."""
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    from graphphysics_b200.synthetic import cylinder_flow_batch
    dev = torch.device("cuda:0")
    b = cylinder_flow_batch(4, seed=3).to(dev)
    torch.manual_seed(0)
    x, Gm = torch.randn(b.x.shape[0], 11, device=dev), torch.randn(b.x.shape[0], 2, device=dev)
    res = {}
    for side in ("0", "1"):
        monkeypatch.setenv("GP_B200_SIDE_REDUCE", side)
        torch.manual_seed(1)
        m = EncodeProcessDecode(4, 11, 3, 2, hidden_size=128).to(dev)
        assert m.engine._side_reduce == (side == "1")
        outs = []
        for _ in range(2):                       # twice: the partial-gradient sets alternate between layers and steps
            out = m(Data(x=x, edge_index=b.edge_index, edge_attr=b.edge_attr))
            (out * Gm).sum().backward()
            outs.append((out.detach().clone(), m.engine.gflat.clone()))
            m.zero_grad(set_to_none=True)
        res[side] = outs
    for (o0, g0), (o1, g1) in zip(res["0"], res["1"]):
        assert torch.equal(o0, o1) and torch.equal(g0, g1)
