"""Full-size (BASELINE.json configs[1]: 32 CylinderFlow-shaped meshes, E = 372 752 directed edges,
H = 128) properties of the CUDA path that need no CPU oracle pass at that size:

* the atomic-free segment sum equals an exact (fp64) index_add of the very bf16 values the kernel
  emitted, for every receiver, including the pieces that the boundary fix-up combines;
* residual identity: with the residual on, output - residual equals the output without it
  (both are bf16 tiles: bit-exact after one bf16 rounding of the sum);
* determinism: two forward+backward passes give bit-identical outputs and gradients (no float
  atomics anywhere);
* a training step under CUDA-graph replay equals the same step launched eagerly, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = {"model": {"type": "epd", "message_passing_num": 3, "hidden_size": 128, "node_input_size": 2, "output_size": 2,
                 "edge_input_size": 3},
       "index": {"feature_index_start": 0, "feature_index_end": 2, "output_index_start": 0, "output_index_end": 2,
                 "node_type_index": 2}}


def _batch(dev):
    from graphphysics_b200.synthetic import cylinder_flow_batch
    return cylinder_flow_batch(32, seed=0).to(dev)


def test_segment_sum_exact_at_full_size():
    from graphphysics_b200 import ops
    from graphphysics_b200.graph import get_csr
    from graphphysics_b200.models.processors import EncodeProcessDecode
    dev = torch.device("cuda:0")
    b = _batch(dev)
    N, E, H = b.x.shape[0], b.edge_index.shape[1], 128
    assert E == 372752
    torch.manual_seed(0)
    eng = EncodeProcessDecode(1, 11, 3, 2, hidden_size=H).to(dev).engine
    g = get_csr(b.edge_index, N)
    bf = torch.bfloat16
    e = torch.randn(E, H, device=dev).to(bf)
    P = torch.randn(N, 3 * H, device=dev).to(bf)
    outs = {}
    for with_resid in (False, True):
        y = torch.empty((E, H), dtype=bf, device=dev)
        agg = torch.full((N, H), float("nan"), device=dev)
        bnd = torch.full((ops.seg_bnd_size(E, H),), float("nan"), device=dev)
        eng._mlp(eng.edge[0], E, e, H, y, H, resid=e if with_resid else None, init=P, init_off0=0, init_off1=H,
                 idx0=g.dst, idx1=g.src, two_inits=True, seg_id=g.dst, seg_out=agg, seg_bnd=bnd)
        ops.seg_fixup(g.rowptr_dst, H, bnd, agg)
        outs[with_resid] = (y, agg)
    torch.cuda.synchronize()
    u, agg = outs[False]                      # no residual: y is exactly the bf16 update the kernel summed
    ref = torch.zeros((N, H), dtype=torch.float64, device=dev).index_add_(0, g.dst.long(), u.double())
    assert torch.isfinite(agg).all()
    err = (agg.double() - ref).abs().max().item()
    assert err <= 1e-5 * max(ref.abs().max().item(), 1.0), err      # fp32 summation order only
    # residual identity on bf16 tiles: y_resid == bf16(e + u), bit for bit
    y_res, agg_res = outs[True]
    assert torch.equal(y_res, (e.float() + u.float()).to(bf))
    assert torch.equal(agg_res, agg)
    # bf16 segment sums (what the engine uses; this call runs the compile-time specialised edge kernel, the
    # fp32 calls above the general one): the same fp32 sums rounded once, and the same outputs, bit for bit
    y16 = torch.empty((E, H), dtype=bf, device=dev)
    agg16 = torch.full((N, H), float("nan"), dtype=bf, device=dev)
    bnd = torch.full((ops.seg_bnd_size(E, H),), float("nan"), device=dev)
    eng._mlp(eng.edge[0], E, e, H, y16, H, resid=e, init=P, init_off0=0, init_off1=H, idx0=g.dst, idx1=g.src,
             two_inits=True, seg_id=g.dst, seg_out=agg16, seg_bnd=bnd)
    ops.seg_fixup(g.rowptr_dst, H, bnd, agg16)
    torch.cuda.synchronize()
    assert torch.equal(y16, y_res)
    assert torch.equal(agg16, agg.to(bf))


def test_forward_backward_is_deterministic_at_full_size():
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    dev = torch.device("cuda:0")
    b = _batch(dev)
    N = b.x.shape[0]
    torch.manual_seed(1)
    model = EncodeProcessDecode(3, 11, 3, 2, hidden_size=128).to(dev)
    x = torch.randn(N, 11, device=dev)
    G = torch.randn(N, 2, device=dev)
    runs = []
    for _ in range(2):
        out = model(Data(x=x, edge_index=b.edge_index, edge_attr=b.edge_attr))
        (out * G).sum().backward()
        torch.cuda.synchronize()
        runs.append((out.detach().clone(), model.engine.gflat.clone()))
    assert torch.equal(runs[0][0], runs[1][0])
    assert torch.equal(runs[0][1], runs[1][1])
    assert torch.isfinite(runs[0][1]).all() and runs[0][1].abs().max() > 0


def test_graph_replay_equals_eager_step_at_full_size():
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")
    b = _batch(dev)
    res = []
    for graphed in (False, True):
        tr = Trainer(CFG, learning_rate=1e-3, num_steps=1000, warmup=10, device=dev, seed=0)
        tr.enable_cuda_graph(graphed)
        losses = [float(tr.training_step(b)) for _ in range(4)]     # graphed: eager step + capture, then 3 replays
        torch.cuda.synchronize()
        res.append((losses, tr.engine.flat.data.clone(), tr.step_index))
    (l_e, p_e, n_e), (l_g, p_g, n_g) = res
    assert n_e == n_g == 4
    assert l_e == l_g, (l_e, l_g)
    assert torch.equal(p_e, p_g)


def test_staged_host_batches_equal_resident_batches():
    """Trainer.stage() (pinned host batch copied on a side stream while the previous step runs) feeds the
    same values as handing a device batch to training_step: identical losses and parameters."""
    from graphphysics_b200.synthetic import cylinder_flow_batch
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")
    host = [cylinder_flow_batch(4, seed=s, pin=True) for s in (0, 1)]
    res = []
    for staged in (False, True):
        tr = Trainer(CFG, learning_rate=1e-3, num_steps=1000, warmup=10, device=dev, seed=0)
        tr.enable_cuda_graph(True)
        losses = []
        if staged:
            tr.stage(host[0])
            for i in range(4):
                loss = tr.training_step(None)
                tr.stage(host[(i + 1) % 2])
                losses.append(float(loss))
        else:
            for i in range(4):
                losses.append(float(tr.training_step(host[i % 2].to(dev))))
        torch.cuda.synchronize()
        res.append((losses, tr.engine.flat.data.clone()))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1])
