"""Goldens of the variant flags from the UNMODIFIED reference (build container only; TEST INFRASTRUCTURE):

    python oracle/make_golden_variants.py      ->  tests/golden/variants.npz

EncodeProcessDecode with use_silu_activation / use_gated_mlp / use_gated_attention (the aggregation gate of
GraphNetBlock, layers.py:1091-1098) / use_rope_embeddings (relative RoPE on the senders, layers.py:1020-1026, 1104-1149)
and EncodeTransformDecode with use_gated_attention / use_rope_embeddings / SiLU gating (layers.py:637-697, 213-249),
both also with use_temporal_block (TemporalAttention, layers.py:822-887):
inputs, weights, output and the gradients of sum(out * G) for every parameter, on the small mesh of the other goldens."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

EPD_CASES = {
    "epd_silu": dict(silu=True, kw={}),
    "epd_gated_mlp": dict(silu=False, kw=dict(use_gated_mlp=True)),
    "epd_gated_mlp_silu": dict(silu=True, kw=dict(use_gated_mlp=True)),
    "epd_gate": dict(silu=False, kw=dict(use_gated_attention=True)),
    "epd_rope": dict(silu=False, kw=dict(use_rope_embeddings=True, rope_pos_dimension=2)),
    "epd_all": dict(silu=True, kw=dict(use_gated_mlp=True, use_gated_attention=True, use_rope_embeddings=True, rope_pos_dimension=2)),
    "epd_temporal": dict(silu=False, kw=dict(use_temporal_block=True)),
}
ETD_CASES = {
    "etd_gated_attention": dict(silu=False, kw=dict(use_gated_attention=True)),
    "etd_rope": dict(silu=False, kw=dict(use_rope_embeddings=True, rope_pos_dimension=2)),
    "etd_silu": dict(silu=True, kw={}),
    "etd_shared_qkv": dict(silu=False, kw=dict(use_separate_proj_weight=False)),
    "etd_temporal": dict(silu=False, kw=dict(use_temporal_block=True)),
}


def main():
    ref = ref_shim.import_reference()
    layers, processors = ref["layers"], ref["processors"]
    from torch_geometric.data import Data
    torch.set_num_threads(4)
    pos, tris = O.grid_tri_mesh(14, 9, jitter=0.3, seed=3, hole=(0.4, 0.2, 0.08))
    ei = O.face_to_edge(tris, len(pos))
    ea = O.edge_features(pos, ei)
    N = len(pos)
    ei_t, ea_t, pos_t = torch.from_numpy(ei), torch.from_numpy(ea), torch.from_numpy(pos)
    store = dict(pos=pos, edge_index=ei, edge_attr=ea)
    torch.manual_seed(11)
    x_epd, g_epd = torch.randn(N, 11), torch.randn(N, 2)
    x_etd, g_etd = torch.randn(N, 23), torch.randn(N, 3)
    phi = torch.rand(N)
    store.update(x_epd=x_epd.numpy(), G_epd=g_epd.numpy(), x_etd=x_etd.numpy(), G_etd=g_etd.numpy(), phi=phi.numpy())
    for name, case in EPD_CASES.items():
        layers.set_use_silu_activation(case["silu"])
        torch.manual_seed(7)
        m = processors.EncodeProcessDecode(2, 11, 3, 2, hidden_size=32, **case["kw"])
        with torch.no_grad():                       # the gate position vector starts at zero: make phi matter
            for n_, p in m.named_parameters():
                if n_.endswith("gate_pos"):
                    p.normal_(0, 0.5)
        g = Data(x=x_epd, edge_index=ei_t, edge_attr=ea_t, pos=pos_t, phi=phi)
        out = m(g)
        (out * g_epd).sum().backward()
        store[f"{name}/out"] = out.detach().numpy()
        for k, v in m.state_dict().items():
            store[f"{name}/sd/{k}"] = v.detach().numpy()
        for k, p in m.named_parameters():
            store[f"{name}/grad/{k}"] = p.grad.numpy()
    for name, case in ETD_CASES.items():
        layers.set_use_silu_activation(case["silu"])
        torch.manual_seed(8)
        m = processors.EncodeTransformDecode(2, 23, 3, hidden_size=64, num_heads=4, **case["kw"])
        out = m(Data(x=x_etd, edge_index=ei_t, pos=pos_t))
        (out * g_etd).sum().backward()
        if name in ("etd_gated_attention", "etd_rope"):
            # return_attention=True of the first block (layers.py:795-801): the values of the sparse softmax, in edge order
            import dgl.sparse as dglsp
            with torch.no_grad():
                h0 = m.nodes_encoder(x_etd)
                adj = dglsp.spmatrix(indices=ei_t, shape=(N, N))
                _, attn = m.processor_list[0](h0, adj, pos=pos_t, return_attention=True)
            store[f"{name}/h0"] = h0.numpy()
            store[f"{name}/attn0"] = attn.val.numpy()
        store[f"{name}/out"] = out.detach().numpy()
        for k, v in m.state_dict().items():
            store[f"{name}/sd/{k}"] = v.detach().numpy()
        for k, p in m.named_parameters():
            store[f"{name}/grad/{k}"] = p.grad.numpy()
    layers.set_use_silu_activation(False)
    np.savez_compressed(f"{OUT}/variants.npz", **store)
    print("wrote", f"{OUT}/variants.npz", len(store), "arrays")


if __name__ == "__main__":
    main()
