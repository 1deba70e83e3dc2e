"""GPU parity at the BENCHMARKED configuration (BASELINE configs[1]: EncodeProcessDecode, 15 message-passing
layers, hidden 128) on one graph of the benchmark batch, and for training_config/cylinder.json verbatim.

Evidence levels and why each bound is what it is:

  * teacher-forced, every one of the 15 layers: each GraphNetBlock is run on the kernel-spec oracle's own layer
    inputs; l2 <= 1e-3 (north-star tolerance) -- what ONE layer of the specialised H=128 kernels adds;
  * free-running output / scalar / gradients vs the kernel-spec oracle: the bf16 specification is chaotic at this
    depth -- the SAME specification evaluated with fp32 instead of fp64 accumulation differs from itself by
    4.6e-2 (output) and 1.8e-1 (gradients) at default init (measured in-test as `floor`), because a one-ulp bf16
    flip is amplified ~5x per layer by the following RMSNorms.  A kernel can therefore only be held to that floor:
    error <= 2 x floor.  A missing gradient term or a wrong sign moves the gradient by O(1) and still fails;
  * vs the fp32 REFERENCE golden (tests/golden/epd_l15_h128.npz): drift of the bf16 format, bounded by 2 x the
    drift of the kernel-spec oracle itself; the reference-tolerance (1e-3) comparison is the tight-mode test
    (tests/test_tight_gpu.py)."""
import json
import os

import numpy as np
import pytest
import torch

from tests.util import check_close, l2_rel

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _bench_case():
    from oracle.cpu_train import default_state_dict
    z = np.load(os.path.join(G, "epd_l15_h128.npz"))
    sd = default_state_dict(15, 11, 3, 2, 128, seed=0)
    for k, v in sd.items():
        got = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        assert np.allclose(got, z["sdsum/" + k], rtol=1e-9, atol=1e-12), f"default init of {k} changed: regenerate the golden"
    return z, sd


def _flat(d, keys):
    return torch.cat([d[k].detach().double().reshape(-1).cpu() for k in keys])


def test_l15_h128_teacher_forced_all_layers():
    from oracle import gp_oracle as O
    from graphphysics_b200 import ops
    from graphphysics_b200.graph import get_csr
    from graphphysics_b200.models.processors import EncodeProcessDecode
    torch.set_num_threads(8)
    dev = torch.device("cuda:0")
    z, sd = _bench_case()
    x, ea, ei = torch.from_numpy(z["x"]), torch.from_numpy(z["edge_attr"]), torch.from_numpy(z["edge_index"])
    N, E, H, L = x.shape[0], ei.shape[1], 128, 15
    model = EncodeProcessDecode(L, 11, 3, 2, hidden_size=H)
    model.load_state_dict(sd)
    model = model.to(dev)
    eng = model.engine
    g = get_csr(ei.to(dev), N)
    perm = g.perm_dst64.cpu()
    sdd = {k: v.double() for k, v in sd.items()}
    rep, ok = [], True
    with torch.no_grad():
        xo = O.rnd(O.mlp(x.double(), sdd, "nodes_encoder", mode="bf16"), "bf16")
        eo = O.rnd(O.mlp(ea.double(), sdd, "edges_encoder", mode="bf16"), "bf16")
        bnd = torch.empty(ops.seg_bnd_size(E, H), dtype=torch.float32, device=dev)
        eng.refresh_weights()
        for l in range(L):
            xn, en = O.graph_net_block(xo, eo, ei[0], ei[1], sdd, f"processor_list.{l}", mode="bf16")
            xk, ek, _ = eng.run_block(l, xo.to(dev).to(torch.bfloat16), eo[perm].to(dev).to(torch.bfloat16).contiguous(),
                                      g, bnd, False)
            ok &= check_close(xk.float(), xn, f"layer {l:2d} x (teacher-forced)", 1e-3, 2e-2, rep)
            ok &= check_close(ek.float(), en[perm], f"layer {l:2d} e (teacher-forced)", 1e-3, 2e-2, rep)
            xo, eo = xn, en
    print("\n".join(rep))
    assert ok, "\n".join(r for r in rep if r.startswith("BAD"))


def test_l15_h128_free_running_output_scalar_gradients():
    from oracle import gp_oracle as O
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    torch.set_num_threads(8)
    dev = torch.device("cuda:0")
    z, sd = _bench_case()
    x, ea, ei, Gm = (torch.from_numpy(z[k]) for k in ("x", "edge_attr", "edge_index", "G"))
    L, H = 15, 128

    def oracle(dt, mode):
        p = {k: v.to(dt).clone().requires_grad_(True) for k, v in sd.items()}
        out = O.epd_forward(p, x.to(dt), ea.to(dt), ei, L, mode=mode)
        s = (out * Gm.to(dt)).sum()
        s.backward()
        return out.detach().double(), float(s), {k: v.grad.double() for k, v in p.items()}

    o64, s64, g64 = oracle(torch.float64, "bf16")        # kernel specification, fp64 accumulate
    o32, s32, g32 = oracle(torch.float32, "bf16")        # the same specification, fp32 accumulate: the noise floor
    model = EncodeProcessDecode(L, 11, 3, 2, hidden_size=H)
    model.load_state_dict(sd)
    model = model.to(dev)
    out = model(Data(x=x.to(dev), edge_index=ei.to(dev), edge_attr=ea.to(dev)))
    s = (out * Gm.to(dev)).sum()
    s.backward()
    torch.cuda.synchronize()
    gk = {k: v.detach().double().cpu() for k, v in model.engine.grads_by_name().items()}
    keys = list(sd.keys())
    floor_out, floor_g = l2_rel(o32, o64), l2_rel(_flat(g32, keys), _flat(g64, keys))
    err_out, err_g = l2_rel(out, o64), l2_rel(_flat(gk, keys), _flat(g64, keys))
    ref_out = torch.from_numpy(z["out"]).double()
    drift_spec, drift_k = l2_rel(o64, ref_out), l2_rel(out, ref_out)
    print(f"output   : kernel vs spec {err_out:.3e}   spec noise floor (fp32 vs fp64 accumulate) {floor_out:.3e}")
    print(f"gradients: kernel vs spec {err_g:.3e}   spec noise floor {floor_g:.3e}")
    print(f"scalar   : kernel {float(s):.5f}  spec {s64:.5f} / {s32:.5f}  reference {float(z['scalar']):.5f}")
    print(f"drift vs fp32 reference golden: spec {drift_spec:.3e}  kernel {drift_k:.3e}")
    assert err_out <= 2.0 * floor_out + 1e-3
    assert err_g <= 2.0 * floor_g + 1e-3
    assert drift_k <= 2.0 * drift_spec + 1e-3
    scale = float(torch.from_numpy(z["G"]).double().norm() * o64.norm())       # |<out, G>| <= |out| |G|
    assert abs(float(s) - s64) <= (2.0 * floor_out + 1e-3) * scale
    # per-tensor: no gradient may be missing or off by a factor (norm ratio to the spec, and to the reference's norms)
    bad = []
    for k in keys:
        r = float(gk[k].norm() / g64[k].norm().clamp_min(1e-30))
        rr = float(gk[k].norm() / max(float(z["gnorm/" + k]), 1e-30))
        if not (0.6 < r < 1.6 and 0.5 < rr < 2.0):
            bad.append((k, r, rr))
    assert not bad, bad[:10]


def test_cylinder_json_verbatim_training_and_prediction():
    """training_config/cylinder.json (epd, 5 layers, hidden 32 -- the general kernels) on the reference's mock cylinder
    trajectory: the reference's own three training steps and eval prediction (tests/golden/cylinder_json_step.npz)."""
    from oracle.cpu_train import CpuTrainer
    from graphphysics_b200.graph import Data
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")
    cfg = json.load(open(os.path.join(G, "training_configs.json")))["cylinder"]
    z = np.load(os.path.join(G, "cylinder_json_step.npz"))
    sd0 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0/")}
    tr = Trainer(cfg, learning_rate=1e-3, num_steps=10, warmup=2, device=dev, seed=0)
    tr.processor.load_state_dict(sd0)
    m, index = cfg["model"], cfg["index"]
    spec = CpuTrainer({k: v.double() for k, v in sd0.items()}, m["message_passing_num"], index, m["output_size"],
                      m["node_input_size"] + 9, m["edge_input_size"], lr=1e-3, num_steps=10, warmup=2, mode="bf16")
    ei, ea, pos = torch.from_numpy(z["edge_index"]), torch.from_numpy(z["edge_attr"]), torch.from_numpy(z["pos"])
    frames, ys = torch.from_numpy(z["frames"]), torch.from_numpy(z["ys"])
    for s in range(3):
        b = Data(x=frames[s].clone(), y=ys[s], pos=pos, edge_index=ei, edge_attr=ea).to(dev)
        loss = float(tr.training_step(b))
        lspec = spec.training_step(frames[s].double(), ys[s].double(), ea.double(), ei)
        print(f"step {s}: loss kernel {loss:.6f}  kernel-spec oracle {lspec:.6f}  reference {z['losses'][s]:.6f}  lr {tr.current_lr():.3e}")
        assert loss == pytest.approx(lspec, rel=3e-3), (s, loss, lspec)                 # same arithmetic: tight
        assert loss == pytest.approx(float(z["losses"][s]), rel=1e-2), (s, loss)        # bf16 drift vs the fp32 reference
        assert tr.current_lr() == pytest.approx(float(z["lrs"][s]), rel=1e-9)
    tr.model.eval()
    with torch.no_grad():
        b = Data(x=frames[3].clone(), y=ys[3], pos=pos, edge_index=ei, edge_attr=ea).to(dev)
        net, tgt, outp = tr.model(b)
    assert l2_rel(tgt, torch.from_numpy(z["eval_target"])) < 1e-4
    assert l2_rel(outp, torch.from_numpy(z["eval_outputs"])) < 1e-2                     # one-step prediction, physical units
    assert l2_rel(net, torch.from_numpy(z["eval_net"])) < 5e-2
