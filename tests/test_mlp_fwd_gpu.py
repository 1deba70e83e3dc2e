"""GPU parity of the fused tcgen05 MLP forward (gp_mlp_fwd) against the CPU oracle in
kernel-arithmetic mode (bf16 MMA operands, fp32/fp64 accumulate).  Tolerance: norm-relative
1e-3 for fp32 outputs, 1 bf16 ulp-ish (4e-3 of max) for bf16-stored outputs."""
import numpy as np
import pytest
import torch

from tests.util import bf16_round, random_sorted_graph, rel_err

pytestmark = pytest.mark.gpu


def _make_mlp(in_size, hidden, out_size, seed, norm=True):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    sizes = [(hidden, in_size), (hidden, hidden), (hidden, hidden), (out_size, hidden)]
    for i, (n, k) in enumerate(sizes):
        bound = 1.0 / np.sqrt(k)
        sd[f"m.{2*i}.weight"] = (torch.rand((n, k), generator=g) * 2 - 1) * bound
        sd[f"m.{2*i}.bias"] = (torch.rand((n,), generator=g) * 2 - 1) * bound
    if norm:
        sd["m.7.scale"] = 1.0 + 0.1 * torch.randn((out_size,), generator=g)
    return sd


def _pack(sd, dev, k0_pad=None, n_last_pad=None):
    from graphphysics_b200 import ops
    ws, bs = [], []
    for i in range(4):
        w = sd[f"m.{2*i}.weight"]
        kp = k0_pad if (i == 0 and k0_pad) else None
        npad = n_last_pad if (i == 3 and n_last_pad) else None
        pw = ops.pack_weight(w.to(dev), npad, kp)
        ws.append(pw)
        bs.append(ops.pack_bias(sd[f"m.{2*i}.bias"].to(dev), pw.shape[0], dev))
    return ws, bs


@pytest.mark.parametrize("hidden", [128, 64, 32])
@pytest.mark.parametrize("rows", [1000, 128, 77])
def test_plain_mlp_norm(hidden, rows):
    from oracle import gp_oracle as O
    from graphphysics_b200 import ops
    dev = torch.device("cuda:0")
    sd = _make_mlp(hidden, hidden, hidden, seed=hidden + rows)
    x = bf16_round(torch.randn(rows, hidden, generator=torch.Generator().manual_seed(1)))
    ref = O.mlp(x.double(), {k: v.double() for k, v in sd.items()}, "m", mode="bf16")
    ws, bs = _pack(sd, dev)
    out = torch.empty((rows, hidden), dtype=torch.float32, device=dev)
    ops.mlp_fwd(rows, hidden, ws, bs, a=x.to(dev).to(torch.bfloat16), ka=hidden,
                norm_scale=sd["m.7.scale"].to(dev), out=out, n_valid=hidden)
    torch.cuda.synchronize()
    # a normalised output leaves the kernel through a bf16 tile (it is rounded once, before the
    # residual / segment sum), so fp32 storage still carries bf16 resolution: one ulp of the largest element is up to 2^-7 = 7.8e-3 of it
    assert rel_err(out, bf16_round(ref.float()).double()) < 8e-3


@pytest.mark.parametrize("hidden", [128, 32])
def test_encoder_and_decoder_shapes(hidden):
    """Small K (11 -> padded 16), small N (2 -> padded 16, fp32 out, no norm)."""
    from oracle import gp_oracle as O
    from graphphysics_b200 import ops
    dev = torch.device("cuda:0")
    rows = 531
    sd = _make_mlp(11, hidden, hidden, seed=3)
    x = bf16_round(torch.randn(rows, 11, generator=torch.Generator().manual_seed(2)))
    ref = O.mlp(x.double(), {k: v.double() for k, v in sd.items()}, "m", mode="bf16")
    ws, bs = _pack(sd, dev)
    xp = torch.zeros((rows, 16), dtype=torch.bfloat16, device=dev)
    xp[:, :11] = x.to(dev)
    out = torch.empty((rows, hidden), dtype=torch.bfloat16, device=dev)
    ops.mlp_fwd(rows, hidden, ws, bs, a=xp, ka=16, norm_scale=sd["m.7.scale"].to(dev), out=out, n_valid=hidden)
    torch.cuda.synchronize()
    assert rel_err(out.float(), ref) < 4e-3

    sd = _make_mlp(hidden, hidden, 2, seed=4, norm=False)
    x = bf16_round(torch.randn(rows, hidden, generator=torch.Generator().manual_seed(5)))
    ref = O.mlp(x.double(), {k: v.double() for k, v in sd.items()}, "m", layer_norm=False, mode="bf16")
    ws, bs = _pack(sd, dev)
    out = torch.full((rows, 2), 7.0, dtype=torch.float32, device=dev)
    ops.mlp_fwd(rows, hidden, ws, bs, a=x.to(dev).to(torch.bfloat16), ka=hidden, out=out, n_valid=2)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-3


@pytest.mark.parametrize("hidden", [128, 64])
def test_projection_single_layer(hidden):
    """One bias-free layer with N = 3*hidden (three 128-column accumulator chunks at H=128)."""
    from graphphysics_b200 import ops
    dev = torch.device("cuda:0")
    rows = 300
    g = torch.Generator().manual_seed(9)
    w = torch.randn((3 * hidden, hidden), generator=g) / np.sqrt(hidden)
    x = bf16_round(torch.randn(rows, hidden, generator=g))
    ref = x.double() @ bf16_round(w).double().T
    out = torch.empty((rows, 3 * hidden), dtype=torch.bfloat16, device=dev)
    ops.mlp_fwd(rows, hidden, [ops.pack_weight(w.to(dev))], [None], a=x.to(dev).to(torch.bfloat16), ka=hidden,
                out=out, n_valid=3 * hidden)
    torch.cuda.synchronize()
    assert rel_err(out.float(), ref) < 4e-3


@pytest.mark.parametrize("hidden", [128, 64, 32])
def test_edge_mode_gather_residual_segment_sum(hidden):
    """Edge configuration: gathered two-row pre-activation, residual, receiver-sorted segment
    sum with empty segments, a segment longer than several tiles, and a ragged last tile."""
    from oracle import gp_oracle as O
    from graphphysics_b200 import ops
    dev = torch.device("cuda:0")
    N, E = 300, 3001
    src, dst = random_sorted_graph(N, E, seed=hidden, max_degree_node=17)
    g = torch.Generator().manual_seed(11)
    sd = _make_mlp(hidden, hidden, hidden, seed=6)
    e = bf16_round(torch.randn(E, hidden, generator=g))
    P = bf16_round(torch.randn(N, 3 * hidden, generator=g))
    sd64 = {k: v.double() for k, v in sd.items()}
    pre = (F_linear(e.double(), bf16_round(sd["m.0.weight"]).double())
           + P[dst, :hidden].double() + P[src, hidden:2 * hidden].double())
    upd = O.mlp(None, sd64, "m", mode="bf16", first_pre=pre)
    ref_e = e.double() + bf16_round(upd.float()).double()
    ref_agg = torch.zeros(N, hidden, dtype=torch.float64).index_add_(0, dst, bf16_round(upd.float()).double())

    ws, bs = _pack(sd, dev)
    out = torch.empty((E, hidden), dtype=torch.bfloat16, device=dev)
    agg = torch.full((N, hidden), float("nan"), dtype=torch.float32, device=dev)
    bnd = torch.full((ops.seg_bnd_size(E, hidden),), float("nan"), dtype=torch.float32, device=dev)
    e_dev = e.to(dev).to(torch.bfloat16)
    dst32, src32 = dst.to(dev).int(), src.to(dev).int()
    ops.mlp_fwd(E, hidden, ws, bs, a=e_dev, ka=hidden, init=P.to(dev).to(torch.bfloat16), init_off0=0,
                init_off1=hidden, idx0=dst32, idx1=src32, two_inits=True, norm_scale=sd["m.7.scale"].to(dev),
                resid=e_dev, out=out, n_valid=hidden, seg_id=dst32, seg_out=agg, seg_bnd=bnd)
    rowptr = torch.zeros(N + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=N), 0)
    ops.seg_fixup(rowptr.to(dev), hidden, bnd, agg)
    torch.cuda.synchronize()
    assert rel_err(out.float(), ref_e) < 4e-3
    assert not torch.isnan(agg).any()
    assert rel_err(agg, ref_agg) < 2e-3


def F_linear(x, w):
    return x @ w.T


def test_node_mode_fp32_operand_direct_init():
    from oracle import gp_oracle as O
    from graphphysics_b200 import ops
    dev = torch.device("cuda:0")
    hidden, N = 128, 700
    g = torch.Generator().manual_seed(21)
    sd = _make_mlp(hidden, hidden, hidden, seed=8)
    agg = torch.randn(N, hidden, generator=g) * 3
    x = bf16_round(torch.randn(N, hidden, generator=g))
    P = bf16_round(torch.randn(N, 3 * hidden, generator=g))
    pre = F_linear(bf16_round(agg).double(), bf16_round(sd["m.0.weight"]).double()) + P[:, 2 * hidden:].double()
    upd = O.mlp(None, {k: v.double() for k, v in sd.items()}, "m", mode="bf16", first_pre=pre)
    ref = x.double() + bf16_round(upd.float()).double()
    ws, bs = _pack(sd, dev)
    out = torch.empty((N, hidden), dtype=torch.bfloat16, device=dev)
    h2 = torch.empty((N, hidden), dtype=torch.bfloat16, device=dev)
    x_dev = x.to(dev).to(torch.bfloat16)
    ops.mlp_fwd(N, hidden, ws, bs, a=agg.to(dev), ka=hidden, init=P.to(dev).to(torch.bfloat16), init_off0=2 * hidden,
                norm_scale=sd["m.7.scale"].to(dev), resid=x_dev, out=out, n_valid=hidden, save_h2=h2)
    torch.cuda.synchronize()
    assert rel_err(out.float(), ref) < 4e-3
    assert float(h2.float().abs().max()) > 0 and float(h2.float().min()) >= 0
