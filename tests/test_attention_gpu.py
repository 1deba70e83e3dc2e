"""GPU parity of the CSR masked-attention kernels and the Transformer model built on them.
The attention kernels compute in fp32 (q / k / v read as fp32 or bf16), so they are compared with the fp64
oracle at 1e-5; the model (bf16 MMA operands on the timed path) is compared with the golden output and
gradients of the UNMODIFIED reference (DGL branch restated in oracle/ref_shim.py) at the bf16 drift level
here, at 1e-3 against the kernel-mode oracle and in precision="tight" at 1e-3 against the same golden in
tests/test_dense_gpu.py."""
import os

import numpy as np
import pytest
import torch

from tests.util import l2_rel, rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("hidden,heads", [(64, 4), (32, 4), (128, 4), (64, 1), (64, 16), (128, 2), (32, 8)])
def test_csr_attention_forward_backward(hidden, heads):
    from oracle import gp_oracle as O
    from graphphysics_b200.graph import GraphCSR
    from graphphysics_b200.ops import CSRAttention
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(hidden + heads)
    n, e = 211, 1500
    ei = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)])
    ei[0, ei[0] == 5] = 6                                           # row 5 has no entries -> y = 0
    ei = np.unique(ei, axis=1)                                      # the adjacency has no duplicate entries
    ei_t = torch.from_numpy(ei)
    g = torch.Generator().manual_seed(0)
    q, k, v, dy = (torch.randn(n, hidden, generator=g) for _ in range(4))
    q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
    d = hidden // heads
    y_ref = O.sparse_attention(q64.reshape(n, d, heads), k64.reshape(n, d, heads), v64.reshape(n, d, heads),
                               ei_t[0], ei_t[1], n).reshape(n, hidden)
    y_ref = torch.nan_to_num(y_ref)                                 # empty rows: oracle divides 0/0
    (y_ref * dy.double()).sum().backward()
    csr = GraphCSR(ei_t.to(dev), n)
    qd, kd, vd = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    y = CSRAttention.apply(qd, kd, vd, csr, heads)
    (y * dy.to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert float(y[5].detach().abs().max()) == 0.0
    assert rel_err(y, y_ref) < 1e-5
    for got, ref in ((qd.grad, q64.grad), (kd.grad, k64.grad), (vd.grad, v64.grad)):
        assert rel_err(got, torch.nan_to_num(ref)) < 2e-5
    # bit-reproducible
    y2 = CSRAttention.apply(qd.detach(), kd.detach(), vd.detach(), csr, heads)
    assert torch.equal(y2, y.detach())
    # bf16 q / k / v (the Transformer path): same arithmetic on the rounded values, y stored as bf16
    qb, kb, vb = (t.to(torch.bfloat16) for t in (q, k, v))
    q64, k64, v64 = (t.double().requires_grad_(True) for t in (qb, kb, vb))
    y_ref = torch.nan_to_num(O.sparse_attention(q64.reshape(n, d, heads), k64.reshape(n, d, heads), v64.reshape(n, d, heads),
                                                ei_t[0], ei_t[1], n).reshape(n, hidden))
    (y_ref * dy.double()).sum().backward()
    qd, kd, vd = (t.to(dev).requires_grad_(True) for t in (qb, kb, vb))
    y = CSRAttention.apply(qd, kd, vd, csr, heads)
    (y.float() * dy.to(dev)).sum().backward()
    assert l2_rel(y, y_ref) < (4e-3 if y.dtype == torch.bfloat16 else 1e-5)
    for got, ref in ((qd.grad, q64.grad), (kd.grad, k64.grad), (vd.grad, v64.grad)):
        assert l2_rel(got.float(), torch.nan_to_num(ref)) < 4e-3


def test_transformer_model_against_reference_golden():
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.processors import EncodeTransformDecode
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(G, "transformer_l2_h64.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    m = EncodeTransformDecode(2, 23, 3, hidden_size=64, num_heads=4)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)
    m = m.to(dev)
    out = m(Data(x=torch.from_numpy(z["x"]).to(dev), edge_index=torch.from_numpy(z["edge_index"]).to(dev)))
    assert l2_rel(out, torch.from_numpy(z["out"])) < 1e-2           # bf16 operands vs the fp32 reference
    (out * torch.from_numpy(z["G"]).to(dev)).sum().backward()
    biggest = max(float(np.linalg.norm(z["grad/" + n])) for n, _ in m.named_parameters())
    for name, p in m.named_parameters():
        ref = torch.from_numpy(z["grad/" + name])
        # k_proj.bias has an analytically zero gradient (softmax is shift-invariant): compare on the
        # scale of the largest gradient tensor when the reference gradient itself is ~0
        err = float((p.grad.cpu() - ref).norm()) / max(float(ref.norm()), 1e-4 * biggest)
        assert err < 0.15, (name, err)      # bf16 operands + ReLU gates of encoder / decoder; tight mode: 1e-3 (test_dense_gpu.py)


def test_transformer_training_steps_run_through_trainer():
    """coarse-aneurysm-style config (type 'transformer'): the Trainer drives the model through autograd
    (native CSR attention kernels inside, flat parameter / gradient buffers, own loss + AdamW kernels);
    the loss falls on a repeated batch."""
    from graphphysics_b200.synthetic import cylinder_flow_batch
    from graphphysics_b200.training.loop import Trainer
    dev = torch.device("cuda:0")
    cfg = {"model": {"type": "transformer", "message_passing_num": 2, "hidden_size": 64, "num_heads": 4,
                     "node_input_size": 2, "output_size": 2, "edge_input_size": 0},
           "index": {"feature_index_start": 0, "feature_index_end": 2, "output_index_start": 0, "output_index_end": 2,
                     "node_type_index": 2}}
    tr = Trainer(cfg, learning_rate=2e-3, num_steps=100, warmup=2, device=dev, seed=0)
    batch = cylinder_flow_batch(2, nx=20, ny=10, seed=0).to(dev)
    p0 = tr.engine.flat.data.clone()
    losses = [float(tr.training_step(batch)) for _ in range(8)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert not torch.equal(p0, tr.engine.flat.data)


@pytest.mark.parametrize("rope", [False, True])
def test_return_attention_matches_oracle(rope):
    """Attention / Transformer .forward(..., return_attention=True) (layers.py:680-697, 795-801): the output is the fused
    path's, the attention values (E x heads, caller's edge order) match the oracle's sparse softmax and sum to one per row."""
    from graphphysics_b200.models.layers import Transformer
    from oracle import gp_oracle as O
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    N, H, heads = 300, 64, 4
    rng = np.random.default_rng(0)
    key = np.unique(rng.integers(0, N, 4000) * N + rng.integers(0, N, 4000))
    ei = torch.from_numpy(np.stack([key // N, key % N]))
    ei = ei[:, torch.randperm(ei.shape[1])]                      # not sorted: the values must follow the caller's order
    blk = Transformer(H, H, heads, use_rope_embeddings=rope, pos_dimension=2).to(dev)
    blk.set_precision("tight")
    x = torch.randn(N, H, device=dev)
    pos = torch.rand(N, 2, device=dev)
    with torch.no_grad():
        plain = blk(x, ei.to(dev), pos=pos)
        out, attn = blk(x, ei.to(dev), pos=pos, return_attention=True)
        a_out, a_attn = blk.attention(x, ei.to(dev), pos=pos, return_attention=True)
    assert torch.equal(out, plain)
    assert tuple(attn.val.shape) == (ei.shape[1], heads) and torch.equal(attn.row.cpu(), ei[0]) and torch.equal(attn.col.cpu(), ei[1])
    sums = torch.zeros(N, heads, device=dev).index_add_(0, ei[0].to(dev), attn.val)
    has = torch.zeros(N, device=dev).index_add_(0, ei[0].to(dev), torch.ones(ei.shape[1], device=dev)) > 0
    assert torch.allclose(sums[has], torch.ones_like(sums[has]), atol=1e-5)
    # oracle: the same projections in fp64
    sd = {k: v.detach().double().cpu() for k, v in blk.state_dict().items()}
    xd = x.double().cpu()
    n1 = O.rms_norm(xd, sd["norm1.scale"])
    d = H // heads
    q = O.linear(n1, sd["attention.q_proj.weight"], sd["attention.q_proj.bias"], None).reshape(N, d, heads)
    k = O.linear(n1, sd["attention.k_proj.weight"], sd["attention.k_proj.bias"], None).reshape(N, d, heads)
    if rope:
        q, k = O.rope_nodes(q, k, pos.double().cpu(), sd["attention.rope_inv_freq"])
    ref = O.attention_values(q, k, ei[0], ei[1], N)
    assert float((attn.val.double().cpu() - ref).abs().max()) < 2e-5
    # the bare Attention module normalises nothing: a different matrix, same shape, rows summing to one
    assert tuple(a_attn.val.shape) == (ei.shape[1], heads) and a_out.shape == x.shape
