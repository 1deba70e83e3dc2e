"""GPU: the online normaliser and the Simulator's node-feature assembly on native kernels (gp_normalizer_*, gp_node_features)
against the reference arithmetic of graphphysics/models/layers.py:281-408 and simulator.py:112-143 written in torch fp64 / fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_stats(xs):
    s = sum(x.double().sum(0) for x in xs)
    q = sum((x.double() ** 2).sum(0) for x in xs)
    n = sum(x.shape[0] for x in xs)
    mean = s / n
    std = torch.sqrt(torch.clamp(q / n - mean ** 2, min=0.0))
    return s, q, n, mean, std


@pytest.mark.parametrize("rows,size", [(64424, 11), (372752, 3), (1923, 2), (7, 64), (1, 5)])
def test_normalizer_matches_reference_arithmetic(rows, size):
    from graphphysics_b200.models.layers import Normalizer
    torch.manual_seed(rows + size)
    nrm = Normalizer(size=size, device=DEV)
    xs = [torch.randn(rows, size, device=DEV) * 3 + 1.5, torch.randn(max(rows // 2, 1), size, device=DEV) - 0.7]
    outs = [nrm(x) for x in xs]
    s, q, n, mean, std = _ref_stats(xs)
    assert torch.allclose(nrm._acc_sum[0].double(), s, rtol=2e-6, atol=1e-3)
    assert torch.allclose(nrm._acc_sum_squared[0].double(), q, rtol=2e-6, atol=1e-3)
    assert float(nrm._acc_count) == n and float(nrm._num_accumulations) == 2
    ref = (xs[1].double() - mean) / torch.clamp(std, min=1e-8)
    assert torch.allclose(outs[1].double(), ref, rtol=1e-4, atol=1e-4)
    # same accumulators through the torch formulas of the class (what a CPU tensor takes): identical up to fp32 rounding
    eager = (xs[1] - nrm._mean()) / nrm._std_with_epsilon()
    assert torch.allclose(outs[1], eager, rtol=1e-6, atol=1e-6)
    back = nrm.inverse(outs[1])
    assert torch.allclose(back, xs[1], rtol=1e-5, atol=1e-5)                  # test_layers.py:92-100 of the reference
    frozen = nrm(xs[0], accumulate=False)
    assert float(nrm._num_accumulations) == 2 and torch.allclose(frozen, (xs[0] - nrm._mean()) / nrm._std_with_epsilon(), rtol=1e-6, atol=1e-6)


def test_normalizer_freezes_on_the_device_after_max_accumulations():
    from graphphysics_b200.models.layers import Normalizer
    nrm = Normalizer(size=4, max_accumulations=2, device=DEV)
    x = torch.randn(100, 4, device=DEV)
    for _ in range(2):
        nrm(x)
    before = nrm._acc_sum.clone()
    nrm._host_calls = 0                      # a replayed CUDA graph cannot re-evaluate the host-side test: the device gate must hold
    nrm(x)
    assert torch.equal(nrm._acc_sum, before) and float(nrm._num_accumulations) == 2 and float(nrm._acc_count) == 200


def test_constant_column_uses_std_epsilon():
    from graphphysics_b200.models.layers import Normalizer
    nrm = Normalizer(size=3, device=DEV)
    x = torch.randn(50, 3, device=DEV)
    x[:, 1] = 2.0
    out = nrm(x)
    assert torch.isfinite(out).all() and float(out[:, 1].abs().max()) == 0.0


def test_node_features_equal_slice_and_one_hot():
    import torch.nn.functional as F
    from graphphysics_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(5000, 7, device=DEV)
    x[:, 4] = torch.randint(0, 9, (5000,), device=DEV).float()
    got = ops.node_features(x, 1, 3, 4, 9)
    ref = torch.cat([x[:, 1:3], F.one_hot(x[:, 4].long(), 9).float()], 1)
    assert torch.equal(got, ref)
    view = x[:, :6]                           # strided rows
    assert torch.equal(ops.node_features(view, 0, 2, 4, 9), torch.cat([x[:, 0:2], F.one_hot(x[:, 4].long(), 9).float()], 1))
