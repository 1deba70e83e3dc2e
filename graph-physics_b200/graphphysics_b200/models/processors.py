"""Encode-process-decode model with the reference's constructor, forward signature and
state_dict keys (graphphysics/models/processors.py:57-215), executed by EPDEngine on the
sm_100a kernels.  `graph` is any object with `.x`, `.edge_index`, `.edge_attr` (a PyG Data or
graphphysics_b200.graph.Data)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from ..graph import get_csr
from .layers import GraphNetBlock, TemporalAttention, build_mlp, use_silu_activation


class EncodeProcessDecode(nn.Module):
    def __init__(self, message_passing_num: int, node_input_size: int, edge_input_size: int, output_size: int,
                 hidden_size: int = 128, only_processor: bool = False, use_rope_embeddings: bool = False,
                 use_gated_attention: bool = False, use_gated_mlp: bool = False, rope_pos_dimension: int = 3,
                 rope_base: float = 10000.0, use_temporal_block: bool = False, precision: str = None):
        super().__init__()
        # "bf16": the fused kernels (bf16 MMA operands and storage, fp32 accumulate) -- the timed path;
        # "tight": split-precision GEMMs (bf16 hi+lo, three MMAs per product, fp32 storage), graphphysics_b200/tight.py
        self.precision = precision or os.environ.get("GP_B200_PRECISION", "bf16")
        if self.precision not in ("bf16", "tight"):
            raise ValueError(f"precision must be 'bf16' or 'tight', got {self.precision!r}")
        self.only_processor = only_processor
        self.hidden_size = hidden_size
        self.d = output_size
        self.use_temporal_block = use_temporal_block
        self.use_gated_mlp = use_gated_mlp
        self.use_rope = use_rope_embeddings
        self.use_gate = use_gated_attention
        self.rope_axes, self.rope_base = rope_pos_dimension, rope_base
        # created first, as in processors.py:122-126 (same default initialisation under a seed)
        self.temporal_block = TemporalAttention(hidden_size=hidden_size) if use_temporal_block else None
        if not only_processor:
            self.nodes_encoder = build_mlp(node_input_size, hidden_size, hidden_size)
            self.edges_encoder = build_mlp(edge_input_size, hidden_size, hidden_size)
            self.decode_module = build_mlp(hidden_size, hidden_size, output_size, layer_norm=False)
        self.processor_list = nn.ModuleList([
            GraphNetBlock(hidden_size=hidden_size, use_gated_mlp=use_gated_mlp, use_rope=use_rope_embeddings,
                          rope_axes=rope_pos_dimension, rope_base=rope_base, use_gate=use_gated_attention)
            for _ in range(message_passing_num)])
        self._engine = None
        self.act = "silu" if use_silu_activation() else "relu"
        # any variant flag (or the global SiLU switch) takes the model off the fused kernels onto the general path
        self.variant = bool(use_rope_embeddings or use_gated_attention or use_gated_mlp or use_temporal_block or self.act != "relu")
        if self.use_rope and self.rope_axes not in (2, 3):
            raise ValueError("rope_pos_dimension must be 2 or 3 when use_rope_embeddings=True.")
        for blk in self.processor_list:
            blk.precision = self.precision
        if self.temporal_block is not None:
            self.temporal_block.precision = self.precision

    @property
    def engine(self):
        from ..engine import EPDEngine
        if self._engine is None or not self._engine.is_bound():
            self._engine = EPDEngine(self)
        return self._engine

    def forward(self, graph) -> torch.Tensor:
        if self.variant:
            from ..variants import epd_forward as variant_forward
            return variant_forward(self, graph, self.act)
        if self.precision == "tight":
            from ..tight import epd_forward
            return epd_forward(self, graph)
        from ..engine import BlockFunction, EPDFunction
        eng = self.engine
        x, edge_attr = graph.x, graph.edge_attr
        g = get_csr(graph.edge_index, x.shape[0])
        if torch.is_grad_enabled():
            if self.only_processor:
                return BlockFunction.apply(eng.flat, x, edge_attr, eng, g)[0]
            return EPDFunction.apply(eng.flat, x, edge_attr, eng, g)
        out, _, _ = eng.forward(x, edge_attr, g, save=False)
        return out.float()


class EncodeTransformDecode(nn.Module):
    """Encoder MLP, L graph-Transformer blocks over the mesh adjacency, decoder MLP
    (graphphysics/models/processors.py:218-384, the DGL branch).  Same constructor and state_dict
    keys.  Every dense layer -- encoder, q/k/v/proj, gated MLP, decoder -- is a tensor-core GEMM of libgp_b200.so
    (gp_gemm, graphphysics_b200/dense.py), the adjacency-masked attention its CSR kernels, norms and the GELU gate its
    row-wise kernels: no library GEMM and no eager elementwise math on the path."""

    def __init__(self, message_passing_num: int, node_input_size: int, output_size: int, hidden_size: int = 128,
                 num_heads: int = 4, only_processor: bool = False, use_proj_bias: bool = True,
                 use_separate_proj_weight: bool = True, use_rope_embeddings: bool = False,
                 use_gated_attention: bool = False, rope_pos_dimension: int = 3, rope_base: float = 10000.0,
                 use_temporal_block: bool = False, precision: str = None):
        super().__init__()
        from .layers import Transformer
        self.hidden_size, self.only_processor, self.d = hidden_size, only_processor, output_size
        self.use_rope_embeddings, self.use_gated_attention = use_rope_embeddings, use_gated_attention
        self.use_temporal_block = use_temporal_block
        if not only_processor:
            self.nodes_encoder = build_mlp(node_input_size, hidden_size, hidden_size)
            self.decode_module = build_mlp(hidden_size, hidden_size, output_size, layer_norm=False)
        self.processor_list = nn.ModuleList([
            Transformer(input_dim=hidden_size, output_dim=hidden_size, num_heads=num_heads, use_proj_bias=use_proj_bias,
                        use_separate_proj_weight=use_separate_proj_weight, use_rope_embeddings=use_rope_embeddings,
                        use_gated_attention=use_gated_attention, pos_dimension=rope_pos_dimension, rope_base=rope_base)
            for _ in range(message_passing_num)])
        self.temporal_block = (TemporalAttention(hidden_size=hidden_size, num_heads=num_heads)
                               if use_temporal_block else None)                                            # processors.py:332-336
        # "bf16": bf16 MMA operands, fp32 accumulate and residual stream -- the timed path; "tight": three-term split
        # GEMMs (csrc/gemm.cu, terms = 3) with fp32 tensors, for rtol-1e-3 parity with the fp32 reference
        self.precision = precision or os.environ.get("GP_B200_PRECISION", "bf16")
        for blk in self.processor_list:
            blk.set_precision(self.precision)
        if self.temporal_block is not None:
            self.temporal_block.precision = self.precision
        self.act = "silu" if use_silu_activation() else "relu"

    def forward(self, graph) -> torch.Tensor:
        x = graph.x
        if x.device.type != "cuda":
            raise RuntimeError("graphphysics_b200 runs on CUDA devices only (there is no CPU fallback)")
        from .. import dense
        terms = 3 if self.precision == "tight" else 1
        g = get_csr(graph.edge_index, x.shape[0])          # processors.py:366: rows edge_index[0], cols edge_index[1]
        pos = getattr(graph, "pos", None)
        if self.use_rope_embeddings and pos is None:
            raise ValueError("use_rope_embeddings=True requires 'pos' attribute in the input graph.")
        if self.act != "relu":                                   # SiLU encoder / decoder: the general MLP path
            from .. import variants
            enc = lambda seq, t: variants.mlp_seq(seq, t, self.act, terms)
        else:
            enc = lambda seq, t: dense.mlp4(seq, t, terms=terms)
        if not self.only_processor:
            x = enc(self.nodes_encoder, x.float())
        prev_x = x
        for block in self.processor_list:
            prev_x = x
            x = block(x, g, pos=pos)
        if self.temporal_block is not None:                      # processors.py:376-377
            x = self.temporal_block(prev_x, x, g)
        return x if self.only_processor else enc(self.decode_module, x)
