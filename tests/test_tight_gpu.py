"""GPU: the tight arithmetic mode (operands split into three bf16 terms, six tcgen05 MMAs per product, fp32 storage;
csrc/gemm3.cu, graphphysics_b200/tight.py) against the fp32 REFERENCE goldens -- the north-star tolerance rtol 1e-3
for one-step predictions, loss and gradients, end to end, including the benchmarked depth and width (15 layers,
hidden 128).

One qualification, measured and stored in the golden (oracle/make_golden_bench.py): at 15 layers the reference's OWN
fp32 gradients differ from the same modules evaluated in fp64 by 8e-4 (median) to 5.6e-3 (max over tensors) --
ReLU gates whose pre-activation lies within fp32 rounding of zero flip.  A gradient tensor is therefore held to
max(1e-3, 3 x that tensor's fp32-vs-fp64 distance `gnoise`): the reference itself is not reproducible below it."""
import json
import os

import numpy as np
import pytest
import torch

from tests.util import l2_rel, rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3          # BASELINE.json north star


@pytest.mark.parametrize("M,N,K,split", [(128, 128, 64, 1), (300, 96, 200, 1), (77, 2, 128, 1), (130, 384, 128, 1),
                                        (128, 128, 5000, 7), (40, 11, 1000, 3)])
def test_gemm3_matches_fp64(M, N, K, split):
    """All three operand-stride patterns the model uses (forward: both K-contiguous; dgrad: B transposed; wgrad: both
    row-contiguous), ragged shapes, split-K; vs an fp64 product: <= 2e-5 (operands carry 16 mantissa bits)."""
    from graphphysics_b200.tight import gemm3
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N)
    for pattern in ("nt", "nn", "tn"):
        A = torch.randn(M, K, generator=g)
        B = torch.randn(N, K, generator=g)
        bias = torch.randn(N, generator=g)
        ref = A.double() @ B.double().t() + bias.double()
        a_dev = (A if pattern != "tn" else A.t().contiguous()).to(dev)            # tn: stored [K, M]
        b_dev = (B if pattern == "nt" else B.t().contiguous()).to(dev)            # nn / tn: stored [K, N]
        a_sm, a_sk = (K, 1) if pattern != "tn" else (1, M)
        b_sn, b_sk = (K, 1) if pattern == "nt" else (1, N)
        c = torch.full((M, N), float("nan"), device=dev)
        gemm3(M, N, K, a_dev, a_sm, a_sk, b_dev, b_sn, b_sk, c, N, 1, bias=bias.to(dev), split_k=split)
        assert l2_rel(c, ref) < 1e-5 and rel_err(c, ref) < 2e-5, (pattern, l2_rel(c, ref))
        c2 = c.clone()
        gemm3(M, N, K, a_dev, a_sm, a_sk, b_dev, b_sn, b_sk, c2, N, 1, relu=True, accumulate=True, split_k=split)
        ref2 = torch.relu(ref + (ref - bias.double()))
        assert l2_rel(c2, ref2) < 1e-5, pattern


def _grads_close(model, z, rep):
    ok = True
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        noise = float(z["gnoise/" + k]) if "gnoise/" + k in z.files else 0.0
        tol = max(TOL, 3.0 * noise)
        if "grad/" + k in z.files:
            ref = torch.from_numpy(z["grad/" + k])
            err = l2_rel(p.grad, ref)
            line = f"  grad {k:45s} l2_rel vs fp32 reference {err:.2e} (tol {tol:.1e})"
            if "grad64/" + k in z.files:
                line += f"; vs fp64 reference {l2_rel(p.grad, torch.from_numpy(z['grad64/' + k])):.2e} (the fp32 reference itself: {noise:.2e})"
            rep.append(("ok  " if err <= tol else "BAD ") + line)
            ok &= err <= tol
        if "gnorm/" + k in z.files:
            gn = float(z["gnorm/" + k])
            ok &= abs(float(p.grad.double().norm()) - gn) <= tol * gn + 1e-9
    return ok


@pytest.mark.parametrize("name", ["epd_l2_h32.npz", "epd_l2_h64.npz"])
def test_tight_epd_matches_reference_golden(name):
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(G, name))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model = EncodeProcessDecode(int(z["L"]), 11, 3, 2, hidden_size=int(z["H"]), precision="tight")
    model.load_state_dict(sd)
    model = model.to(dev)
    out = model(Data(x=torch.from_numpy(z["x"]).to(dev), edge_index=torch.from_numpy(z["edge_index"]).to(dev),
                     edge_attr=torch.from_numpy(z["edge_attr"]).to(dev)))
    (out * torch.from_numpy(z["G"]).to(dev)).sum().backward()
    rep = [f"{name}: output l2_rel {l2_rel(out, torch.from_numpy(z['out'])):.2e}"]
    ok = l2_rel(out, torch.from_numpy(z["out"])) <= TOL and rel_err(out, torch.from_numpy(z["out"])) <= TOL
    ok &= _grads_close(model, z, rep)
    print("\n".join(rep))
    assert ok, "\n".join(rep)


def test_tight_benchmark_config_l15_h128_matches_reference_golden():
    """BASELINE configs[1] (15 layers, hidden 128) on one graph of the benchmark batch: one-step output, scalar loss
    and gradients within rtol 1e-3 of the UNMODIFIED reference (tests/golden/epd_l15_h128.npz)."""
    from oracle.cpu_train import default_state_dict
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeProcessDecode
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(G, "epd_l15_h128.npz"))
    sd = default_state_dict(15, 11, 3, 2, 128, seed=0)
    for k, v in sd.items():
        got = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        assert np.allclose(got, z["sdsum/" + k], rtol=1e-9, atol=1e-12), f"default init of {k} changed: regenerate the golden"
    model = EncodeProcessDecode(15, 11, 3, 2, hidden_size=128, precision="tight")
    model.load_state_dict(sd)
    model = model.to(dev)
    out = model(Data(x=torch.from_numpy(z["x"]).to(dev), edge_index=torch.from_numpy(z["edge_index"]).to(dev),
                     edge_attr=torch.from_numpy(z["edge_attr"]).to(dev)))
    s = (out * torch.from_numpy(z["G"]).to(dev)).sum()
    s.backward()
    ref = torch.from_numpy(z["out"])
    rep = [f"output l2_rel {l2_rel(out, ref):.2e} max_rel {rel_err(out, ref):.2e}; scalar {float(s):.6f} vs {float(z['scalar']):.6f}"]
    ok = l2_rel(out, ref) <= TOL and rel_err(out, ref) <= TOL
    ok &= abs(float(s) - float(z["scalar"])) <= TOL * float(ref.norm() * torch.from_numpy(z["G"]).norm())
    ok &= _grads_close(model, z, rep)
    print("\n".join(rep))
    assert ok, "\n".join(rep)


def test_tight_training_steps_match_reference_golden():
    """cylinder.json verbatim through the Trainer in tight mode: the reference's three losses (rtol 1e-3), the
    parameters after three AdamW steps, and the eval-mode one-step prediction."""
    import copy
    from graphphysics_b200.graph import Data
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")
    cfg = copy.deepcopy(json.load(open(os.path.join(G, "training_configs.json")))["cylinder"])
    cfg["model"]["precision"] = "tight"
    z = np.load(os.path.join(G, "cylinder_json_step.npz"))
    sd0 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0/")}
    tr = Trainer(cfg, learning_rate=1e-3, num_steps=10, warmup=2, device=dev, seed=0)
    assert not tr.fused
    tr.processor.load_state_dict(sd0)
    ei, ea, pos = torch.from_numpy(z["edge_index"]), torch.from_numpy(z["edge_attr"]), torch.from_numpy(z["pos"])
    frames, ys = torch.from_numpy(z["frames"]), torch.from_numpy(z["ys"])
    for s in range(3):
        b = Data(x=frames[s].clone(), y=ys[s], pos=pos, edge_index=ei, edge_attr=ea).to(dev)
        loss = float(tr.training_step(b))
        print(f"step {s}: loss {loss:.7f}  reference {z['losses'][s]:.7f}")
        assert loss == pytest.approx(float(z["losses"][s]), rel=TOL)
    worst = 0.0
    for k, v in tr.model.state_dict().items():
        if "sd3/" + k in z.files and v.dtype.is_floating_point and v.numel() > 1:
            worst = max(worst, l2_rel(v, torch.from_numpy(z["sd3/" + k])))
    print(f"parameters after 3 steps: worst l2_rel {worst:.2e}")
    assert worst <= 5e-3          # Adam's sign-like first steps amplify tiny gradient differences on near-zero entries
    tr.model.eval()
    with torch.no_grad():
        b = Data(x=frames[3].clone(), y=ys[3], pos=pos, edge_index=ei, edge_attr=ea).to(dev)
        net, tgt, outp = tr.model(b)
    assert l2_rel(outp, torch.from_numpy(z["eval_outputs"])) <= TOL
