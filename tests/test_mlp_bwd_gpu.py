"""GPU parity of the two-layer tcgen05 backward stages (gp_mlp_bwd_stage) against autograd
through the CPU oracle in kernel-arithmetic mode (fp64 accumulate, bf16 rounding at the
kernel's rounding points).  Inputs are conditioned away from the ReLU kink (tests/util.py).
Tolerances: norm-relative (l2) 1e-3 for fp32 gradient sums, 3e-3 for gradients stored as bf16
(half an ulp is 2e-3); max-relative 2e-3 / 6e-3."""
import numpy as np
import pytest
import torch

from tests.util import bf16_round, check_close, condition_rows, random_sorted_graph

pytestmark = pytest.mark.gpu


def _make(hidden, k_in, n_out, seed, norm=True):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i, (n, k) in enumerate([(hidden, k_in), (hidden, hidden), (hidden, hidden), (n_out, hidden)]):
        bound = 1.0 / np.sqrt(k)
        sd[f"m.{2*i}.weight"] = (torch.rand((n, k), generator=g) * 2 - 1) * bound
        sd[f"m.{2*i}.bias"] = (torch.rand((n,), generator=g) * 2 - 1) * bound
    if norm:
        sd["m.7.scale"] = 1.0 + 0.1 * torch.randn((n_out,), generator=g)
    return sd


def _collect(ops, partials, grid, hidden, ka, nb, dev):
    """Reduce the per-CTA partial blocks of one stage into dense gradient tensors."""
    o_dwb, o_dwa, o_dbb, o_dba, o_dsc, stride = ops.bwd_layout(hidden, ka, nb)
    dwb = torch.zeros((nb, hidden), device=dev)
    dwa = torch.zeros((hidden, ka), device=dev)
    dbb = torch.zeros((nb,), device=dev)
    dba = torch.zeros((hidden,), device=dev)
    dsc = torch.zeros((hidden,), device=dev)
    ops.reduce_partials(partials, grid, stride, o_dwb, nb, hidden, hidden, dwb, hidden, False)
    ops.reduce_partials(partials, grid, stride, o_dwa, hidden, ka, ka, dwa, ka, False)
    ops.reduce_partials(partials, grid, stride, o_dbb, 1, nb, nb, dbb, nb, False)
    ops.reduce_partials(partials, grid, stride, o_dba, 1, hidden, hidden, dba, hidden, False)
    ops.reduce_partials(partials, grid, stride, o_dsc, 1, hidden, hidden, dsc, hidden, False)
    return dwb, dwa, dbb, dba, dsc


def _conditioned_input(rows, width, sd, seed, first_pre_of=None, layer_norm=True):
    """Rows of bf16-exact N(0,1) inputs such that no ReLU pre-activation of the 4-layer MLP
    `sd` is within 1e-4 of zero."""
    from oracle import gp_oracle as O
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(rows, width)
    sd64 = {k: v.double() for k, v in sd.items()}

    def draw(idx):
        x[idx] = bf16_round(torch.randn(idx.numel(), width, generator=g))

    def pre(idx):
        zs = []
        if first_pre_of is None:
            O.mlp(x[idx].double(), sd64, "m", layer_norm=layer_norm, mode="bf16", preacts=zs)
        else:
            O.mlp(None, sd64, "m", layer_norm=layer_norm, mode="bf16", first_pre=first_pre_of(x[idx].double(), idx),
                  preacts=zs)
        return zs

    condition_rows(draw, pre, rows)
    return x


@pytest.mark.parametrize("hidden", [128, 64, 32])
def test_edge_block_backward(hidden):
    from oracle import gp_oracle as O
    from graphphysics_b200 import ops
    from tests.test_mlp_fwd_gpu import _pack
    dev = torch.device("cuda:0")
    N, E = 257, 2900
    src, dst = random_sorted_graph(N, E, seed=hidden + 1, max_degree_node=5)
    g = torch.Generator().manual_seed(31)
    sd = _make(hidden, hidden, hidden, seed=12)
    P = bf16_round(torch.randn(N, 3 * hidden, generator=g))
    G1 = bf16_round(torch.randn(E, hidden, generator=g))          # dL/de'  (bf16 in HBM)
    G2 = torch.randn(N, hidden, generator=g)                      # dL/dagg (fp32)
    w0 = bf16_round(sd["m.0.weight"]).double()

    def first_pre(e_rows, idx):
        return e_rows @ w0.T + P[dst[idx], :hidden].double() + P[src[idx], hidden:2 * hidden].double()

    e = _conditioned_input(E, hidden, sd, seed=32, first_pre_of=first_pre)

    # ---- oracle (fp64, kernel arithmetic)
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    e64 = e.double().requires_grad_(True)
    P64 = P.double().requires_grad_(True)
    pre = (torch.nn.functional.linear(e64, O.rnd(sd64["m.0.weight"], "bf16"))
           + P64[dst, :hidden] + P64[src, hidden:2 * hidden])
    upd = O.mlp(None, sd64, "m", mode="bf16", first_pre=pre)
    e_new = O.rnd(e64 + O.rnd(upd, "bf16"), "bf16")
    agg = torch.zeros(N, hidden, dtype=torch.float64).index_add_(0, dst, O.grad_rnd(O.rnd(upd, "bf16"), "bf16"))
    ((e_new * G1.double()).sum() + (agg * G2.double()).sum()).backward()

    # ---- kernels
    ws, bs = _pack(sd, dev)
    e_d = e.to(dev).to(torch.bfloat16)
    P_d = P.to(dev).to(torch.bfloat16)
    dst32, src32 = dst.to(dev).int(), src.to(dev).int()
    out = torch.empty((E, hidden), dtype=torch.bfloat16, device=dev)
    h2 = torch.empty((E, hidden), dtype=torch.bfloat16, device=dev)
    aggd = torch.zeros((N, hidden), device=dev)
    bnd = torch.zeros((ops.seg_bnd_size(E, hidden),), device=dev)
    scale = sd["m.7.scale"].to(dev)
    ops.mlp_fwd(E, hidden, ws, bs, a=e_d, ka=hidden, init=P_d, init_off0=0, init_off1=hidden, idx0=dst32, idx1=src32,
                two_inits=True, norm_scale=scale, resid=e_d, out=out, n_valid=hidden, save_h2=h2, seg_id=dst32,
                seg_out=aggd, seg_bnd=bnd)
    stride = ops.bwd_layout(hidden, hidden, hidden)[5]
    part = torch.zeros((ops.sm_count() * stride,), device=dev)
    G1_d, G2_d = G1.to(dev).to(torch.bfloat16), G2.to(dev)
    delta2 = torch.empty((E, hidden), dtype=torch.bfloat16, device=dev)
    gridB = ops.mlp_bwd_stage(E, hidden, a=h2, ka=hidden, wa=ws[2], ba=bs[2], wb=ws[3], bb=bs[3], partials=part,
                              norm_scale=scale, gy=G1_d, gy_gather=G2_d, gy_idx=dst32, out=delta2, mask_by_ain=True)
    dW3, dW2, db3, db2, dsc = _collect(ops, part, gridB, hidden, hidden, hidden, dev)
    dE = torch.empty((E, hidden), dtype=torch.bfloat16, device=dev)
    delta1 = torch.empty((E, hidden), dtype=torch.bfloat16, device=dev)
    dPd = torch.full((N, hidden), float("nan"), device=dev)
    bnd2 = torch.zeros((ops.seg_bnd_size(E, hidden, backward=True),), device=dev)
    gridA = ops.mlp_bwd_stage(E, hidden, a=e_d, ka=hidden, wa=ws[0], ba=bs[0], wb=ws[1], bb=bs[1], partials=part,
                              init=P_d, init_off0=0, init_off1=hidden, idx0=dst32, idx1=src32, two_inits=True,
                              delta_b=delta2, out=dE, out_resid=G1_d, delta_a_out=delta1, seg_id=dst32, seg_out=dPd,
                              seg_bnd=bnd2)
    dW1, dW0, db1, db0, _ = _collect(ops, part, gridA, hidden, hidden, hidden, dev)
    rowptr = torch.zeros(N + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=N), 0)
    ops.seg_fixup(rowptr.to(dev), hidden, bnd2, dPd, backward=True)
    torch.cuda.synchronize()

    rep, ok = [], True
    ok &= check_close(dE.float(), e64.grad, "dE (bf16)", 3e-3, 6e-3, rep)
    ok &= check_close(dPd, P64.grad[:, :hidden], "dPd", 1e-3, 2e-3, rep)
    dPs = torch.zeros(N, hidden, dtype=torch.float64).index_add_(0, src, delta1.float().double().cpu())
    ok &= check_close(dPs, P64.grad[:, hidden:2 * hidden], "dPs (host sum of d1)", 1e-3, 2e-3, rep)
    for name, got in (("m.6.weight", dW3), ("m.4.weight", dW2), ("m.2.weight", dW1), ("m.0.weight", dW0),
                      ("m.6.bias", db3), ("m.4.bias", db2), ("m.2.bias", db1), ("m.0.bias", db0), ("m.7.scale", dsc)):
        ok &= check_close(got, sd64[name].grad, name, 1e-3, 2e-3, rep)
    print("\n".join(rep))
    assert ok, "\n".join(rep)


@pytest.mark.parametrize("hidden", [128, 32])
def test_encoder_and_decoder_backward(hidden):
    """Encoder shape (K=11 padded to 16, no input gradient) and decoder shape (no norm, 2
    outputs padded to 16, delta given)."""
    from oracle import gp_oracle as O
    from graphphysics_b200 import ops
    from tests.test_mlp_fwd_gpu import _pack
    dev = torch.device("cuda:0")
    rows = 700
    g = torch.Generator().manual_seed(41)
    # ---- encoder: stage B (NORM, fp32 upstream) then stage A (GIVEN, ka=16, no d_in)
    sd = _make(hidden, 11, hidden, seed=13)
    x = _conditioned_input(rows, 11, sd, seed=42)
    G = torch.randn(rows, hidden, generator=g)
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    (O.mlp(x.double(), sd64, "m", mode="bf16") * G.double()).sum().backward()
    ws, bs = _pack(sd, dev)
    xp = torch.zeros((rows, 16), dtype=torch.bfloat16, device=dev)
    xp[:, :11] = x.to(dev)
    out = torch.empty((rows, hidden), dtype=torch.bfloat16, device=dev)
    h2 = torch.empty((rows, hidden), dtype=torch.bfloat16, device=dev)
    scale = sd["m.7.scale"].to(dev)
    ops.mlp_fwd(rows, hidden, ws, bs, a=xp, ka=16, norm_scale=scale, out=out, n_valid=hidden, save_h2=h2)
    part = torch.zeros((ops.sm_count() * ops.bwd_layout(hidden, hidden, hidden)[5],), device=dev)
    delta2 = torch.empty((rows, hidden), dtype=torch.bfloat16, device=dev)
    gB = ops.mlp_bwd_stage(rows, hidden, a=h2, ka=hidden, wa=ws[2], ba=bs[2], wb=ws[3], bb=bs[3], partials=part,
                           norm_scale=scale, gy=G.to(dev), out=delta2, mask_by_ain=True)
    dW3, dW2, db3, db2, dsc = _collect(ops, part, gB, hidden, hidden, hidden, dev)
    gA = ops.mlp_bwd_stage(rows, hidden, a=xp, ka=16, wa=ws[0], ba=bs[0], wb=ws[1], bb=bs[1], partials=part,
                           delta_b=delta2)
    dW1, dW0, db1, db0, _ = _collect(ops, part, gA, hidden, 16, hidden, dev)
    torch.cuda.synchronize()
    rep, ok = [], True
    for name, got in (("m.6.weight", dW3), ("m.4.weight", dW2), ("m.2.weight", dW1), ("m.0.weight", dW0[:, :11]),
                      ("m.6.bias", db3), ("m.4.bias", db2), ("m.2.bias", db1), ("m.0.bias", db0), ("m.7.scale", dsc)):
        ok &= check_close(got, sd64[name].grad, "encoder " + name, 1e-3, 2e-3, rep)
    assert ok, "\n".join(rep)

    # ---- decoder: no norm, 2 outputs: stage B is GIVEN with nb=16
    sd = _make(hidden, hidden, 2, seed=14, norm=False)
    x = _conditioned_input(rows, hidden, sd, seed=43, layer_norm=False)
    G = bf16_round(torch.randn(rows, 2, generator=g))
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    (O.mlp(x64, sd64, "m", layer_norm=False, mode="bf16") * G.double()).sum().backward()
    ws, bs = _pack(sd, dev)
    x_d = x.to(dev).to(torch.bfloat16)
    out = torch.empty((rows, 2), dtype=torch.float32, device=dev)
    ops.mlp_fwd(rows, hidden, ws, bs, a=x_d, ka=hidden, out=out, n_valid=2, save_h2=h2)
    Gp = torch.zeros((rows, 16), dtype=torch.bfloat16, device=dev)
    Gp[:, :2] = G.to(dev)
    gB = ops.mlp_bwd_stage(rows, hidden, a=h2, ka=hidden, wa=ws[2], ba=bs[2], wb=ws[3], bb=bs[3], partials=part,
                           delta_b=Gp, out=delta2, mask_by_ain=True)
    dW3, dW2, db3, db2, _ = _collect(ops, part, gB, hidden, hidden, 16, dev)
    dx = torch.empty((rows, hidden), dtype=torch.float32, device=dev)
    gA = ops.mlp_bwd_stage(rows, hidden, a=x_d, ka=hidden, wa=ws[0], ba=bs[0], wb=ws[1], bb=bs[1], partials=part,
                           delta_b=delta2, out=dx)
    dW1, dW0, db1, db0, _ = _collect(ops, part, gA, hidden, hidden, hidden, dev)
    torch.cuda.synchronize()
    rep, ok = [], True
    ok &= check_close(dx, x64.grad, "decoder dx", 1e-3, 2e-3, rep)
    for name, got in (("m.6.weight", dW3[:2]), ("m.4.weight", dW2), ("m.2.weight", dW1), ("m.0.weight", dW0),
                      ("m.6.bias", db3[:2]), ("m.4.bias", db2), ("m.2.bias", db1), ("m.0.bias", db0)):
        ok &= check_close(got, sd64[name].grad, "decoder " + name, 1e-3, 2e-3, rep)
    assert ok, "\n".join(rep)


def test_specialised_backward_instantiations_equal_general(tmp_path):
    """The compile-time specialised stage kernels (edge stage B / A: TMA everywhere, bias gradients folded into
    the weight-gradient MMAs, interleaved tile layout) must give the same bits as the general kernel on the
    same inputs: every output tensor and every weight / bias / scale gradient partial sum."""
    import os
    import subprocess
    import sys
    helper = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cmp_bwd_modes.py")
    for general in (True, False):
        env = dict(os.environ)
        env.pop("GP_BWD_GENERAL", None)
        if general:
            env["GP_BWD_GENERAL"] = "1"
        subprocess.run([sys.executable, helper, str(tmp_path)], check=True, env=env, capture_output=True, timeout=300)
    a, b = torch.load(tmp_path / "cmp_general.pt"), torch.load(tmp_path / "cmp_fast.pt")
    assert set(a) == set(b) and len(a) == 6
    for k in a:
        assert torch.isfinite(b[k]).all(), k
        assert torch.equal(a[k], b[k]), k
