"""Layer classes of the hot path with the reference's constructor signatures and state_dict keys
(graphphysics/models/layers.py), executing on the sm_100a kernels of libgp_b200.so.

Covered here: RMSNorm (73-129), build_mlp (163-210), Normalizer (281-408), GraphNetBlock
(890-1149), GatedMLP / Attention / Transformer (213-278, 564-819).  The default configuration of every
shipped training_config runs on the fused kernels; the variant flags (SiLU, gated MLP, relative RoPE,
aggregation gate, gated attention, RoPE on q / k) run on the general path of
graphphysics_b200/variants.py and dense.py -- the same native kernels, composed per layer; so does
TemporalAttention (822-887, use_temporal_block).
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from ..graph import get_csr

_USE_SILU_ACTIVATION = False
_MEMORY_OPTIMIZED_TRAINING = False


def set_use_silu_activation(use_silu: bool) -> None:
    global _USE_SILU_ACTIVATION
    _USE_SILU_ACTIVATION = bool(use_silu)


def use_silu_activation() -> bool:
    return _USE_SILU_ACTIVATION


def set_memory_optimized_training(enabled: bool) -> None:
    """The reference toggles bf16 autocast + activation checkpointing here (layers.py:24-36, 803-814).  The
    arithmetic here is bf16 operands / fp32 accumulate either way; the flag switches the encode-process-decode
    engine to checkpointed training: only the inputs of every 4th message-passing layer are kept, the backward
    re-runs one segment's forward at a time (engine.checkpoint_every) -- about 40 % of the activation memory for
    one extra forward pass, bit-identical gradients."""
    global _MEMORY_OPTIMIZED_TRAINING
    _MEMORY_OPTIMIZED_TRAINING = bool(enabled)


def use_memory_optimized_training() -> bool:
    return _MEMORY_OPTIMIZED_TRAINING


class RMSNorm(nn.Module):
    """scale * x / (||x||_2 / sqrt(d) + eps): eps is added to the RMS, not under the root
    (layers.py:104-129).  Inside the fused kernels this is the epilogue of the last MLP layer;
    called on its own it is a small elementwise op on the caller's device."""

    def __init__(self, d: int, p: float = -1.0, eps: float = 1e-8, bias: bool = False):
        super().__init__()
        if not (p < 0.0 or p > 1.0):
            raise NotImplementedError("partial RMSNorm (0 <= p <= 1) is not part of the accelerated path")
        if bias:
            raise NotImplementedError("RMSNorm(bias=True) is not part of the accelerated path")
        self.d, self.p, self.eps, self.bias = d, p, eps, bias
        self.scale = nn.Parameter(torch.ones(d))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.is_cuda and self.d in (32, 64, 128) and self.eps == 1e-8:      # gp_rmsnorm_fwd / _bwd
            from .. import dense
            return dense.rms_norm(x.reshape(-1, self.d), self.scale).reshape(x.shape)
        rms = x.norm(2, dim=-1, keepdim=True) / math.sqrt(self.d)
        return self.scale * (x / (rms + self.eps))


class MLP(nn.Sequential):
    """Container with the reference's Sequential layout (Linear at 0,2,4,6; RMSNorm at 7) so that
    state_dict keys match.  The encode-process-decode engine reads its parameters and runs the
    fused kernels; `forward` on the bare container is the same arithmetic spelled with torch ops
    and exists for shape tests and odd sizes only -- it is not on any timed path."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # noqa: D401
        for m in self:
            x = m(x)
        return x


def build_mlp(in_size: int, hidden_size: int, out_size: int, nb_of_layers: int = 4, layer_norm: bool = True,
              act: Optional[str] = None) -> nn.Module:
    """Linear, act, [Linear, act] x (nb-2), Linear, [RMSNorm]  (layers.py:163-210)."""
    assert nb_of_layers >= 2, "The MLP must have at least 2 layers (input and output)."
    key = act if act is not None else ("silu" if _USE_SILU_ACTIVATION else "relu")
    acts = {"relu": nn.ReLU, "gelu": nn.GELU, "silu": nn.SiLU}
    if key not in acts:
        raise NotImplementedError(f"Activation '{key}' not supported. Available: {list(acts)}.")
    layers = [nn.Linear(in_size, hidden_size), acts[key]()]
    for _ in range(nb_of_layers - 2):
        layers += [nn.Linear(hidden_size, hidden_size), acts[key]()]
    layers.append(nn.Linear(hidden_size, out_size))
    if layer_norm:
        layers.append(RMSNorm(out_size))
    return MLP(*layers)


class Normalizer(nn.Module):
    """Online feature normaliser (layers.py:281-408): running sum, sum of squares and count in
    fp32 buffers (same names, so checkpoints load 1:1), frozen after `max_accumulations` calls.
    The call counter is mirrored on the host so the `if` of layers.py:347 costs no device sync."""

    def __init__(self, size: int, max_accumulations: int = 10 ** 5, std_epsilon: float = 1e-8, name: str = "Normalizer",
                 device: Optional[Union[str, torch.device]] = "cuda"):
        super().__init__()
        self.name, self.device = name, device
        self._max_accumulations = max_accumulations
        self._std_epsilon = torch.tensor(std_epsilon, dtype=torch.float32, device=device)
        self.register_buffer("_acc_count", torch.tensor(0.0, device=device))
        self.register_buffer("_num_accumulations", torch.tensor(0.0, device=device))
        self.register_buffer("_acc_sum", torch.zeros((1, size), dtype=torch.float32, device=device))
        self.register_buffer("_acc_sum_squared", torch.zeros((1, size), dtype=torch.float32, device=device))
        self._host_calls: Optional[int] = 0
        self._eps_host = float(std_epsilon)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._host_calls = None          # re-read the device counter once after a load

    def _native(self, t: torch.Tensor) -> bool:
        """CUDA fp32 rows that need no gradient: the normaliser kernels of libgp_b200.so (three launches instead of ~20)."""
        return (t.is_cuda and self._acc_sum.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1
                and 1 <= t.shape[1] <= 64 and t.shape[1] == self._acc_sum.shape[1] and not (t.requires_grad and torch.is_grad_enabled()))

    def forward(self, batched_data: torch.Tensor, accumulate: bool = True) -> torch.Tensor:
        native = self._native(batched_data)
        if accumulate:
            if self._host_calls is None:
                self._host_calls = int(self._num_accumulations.item())
            if self._host_calls < self._max_accumulations:
                if native:
                    from .. import ops
                    ops.normalizer_update(ops.normalizer_stats(batched_data), batched_data.shape[0], norm=self)
                    self._host_calls += 1
                else:
                    self._accumulate(batched_data.detach())
        if native:
            from .. import ops
            return ops.normalizer_apply(self, batched_data)
        return (batched_data - self._mean()) / self._std_with_epsilon()

    def inverse(self, normalized_batch_data: torch.Tensor) -> torch.Tensor:
        if self._native(normalized_batch_data):
            from .. import ops
            return ops.normalizer_apply(self, normalized_batch_data, inverse=True)
        return normalized_batch_data * self._std_with_epsilon() + self._mean()

    def _accumulate(self, batched_data: torch.Tensor):
        # `gate` repeats the freeze test of layers.py:347 on the device, so a captured CUDA graph of the
        # step (which cannot re-evaluate the host-side `if`) stops accumulating at the same call
        gate = (self._num_accumulations < self._max_accumulations).to(torch.float32)
        self._acc_sum += gate * batched_data.sum(dim=0, keepdim=True)
        self._acc_sum_squared += gate * (batched_data ** 2).sum(dim=0, keepdim=True)
        self._acc_count += gate * batched_data.shape[0]
        self._num_accumulations += gate
        self._host_calls += 1

    def _mean(self) -> torch.Tensor:
        return self._acc_sum / torch.clamp(self._acc_count, min=1.0)

    def _std_with_epsilon(self) -> torch.Tensor:
        var = self._acc_sum_squared / torch.clamp(self._acc_count, min=1.0) - self._mean() ** 2
        return torch.maximum(torch.sqrt(torch.clamp(var, min=0.0)), self._std_epsilon.to(var.device))

    def get_variable(self) -> Dict[str, Any]:
        return {"_max_accumulations": self._max_accumulations, "_std_epsilon": self._std_epsilon,
                "_acc_count": self._acc_count, "_num_accumulations": self._num_accumulations,
                "_acc_sum": self._acc_sum, "_acc_sum_squared": self._acc_sum_squared, "name": self.name}


class _ProcessorStack(nn.Module):
    """Adapter that lets the engine drive a bare list of GraphNetBlocks (latent in, latent out)."""

    def __init__(self, blocks, hidden_size: int):
        super().__init__()
        self.processor_list = nn.ModuleList(blocks)
        self.hidden_size = hidden_size
        self.only_processor = True


class GraphNetBlock(nn.Module):
    """One message-passing step (layers.py:890-1149):
        e' = e + MLP_e([e, x[dst], rope(x[src])]);  agg[n] = gate(x) * sum_{dst=n} (e'-e);  x' = x + MLP_n([x, agg]).
    Same constructor and state_dict keys as the reference (edge_block.*, node_block.*, gate_proj.*, gate_pos).  The default
    flags run on the fused kernels; use_rope / use_gated_mlp / use_gate / the global SiLU switch on the general path of
    graphphysics_b200/variants.py."""

    def __init__(self, hidden_size: int, nb_of_layers: int = 4, layer_norm: bool = True, use_rope: bool = False,
                 rope_axes: int = 3, rope_base: float = 10000.0, use_gated_mlp: bool = False, use_gate: bool = False):
        super().__init__()
        self.hidden_size = hidden_size
        self.use_gated_mlp, self.use_rope, self.use_gate = use_gated_mlp, use_rope, use_gate
        self.rope_axes, self.rope_base = rope_axes, rope_base
        self.act = "silu" if use_silu_activation() else "relu"
        # the fused kernels implement the 4-layer, RMS-normalised MLP every shipped configuration uses; any other depth, or
        # no norm, takes the general path like the flags do
        self.variant = use_rope or use_gated_mlp or use_gate or self.act != "relu" or nb_of_layers != 4 or not layer_norm
        if use_gated_mlp:
            self.edge_block = build_gated_mlp(in_size=3 * hidden_size, hidden_size=hidden_size, out_size=hidden_size)
            self.node_block = build_gated_mlp(in_size=2 * hidden_size, hidden_size=hidden_size, out_size=hidden_size)
        else:
            self.edge_block = build_mlp(3 * hidden_size, hidden_size, hidden_size, nb_of_layers, layer_norm)
            self.node_block = build_mlp(2 * hidden_size, hidden_size, hidden_size, nb_of_layers, layer_norm)
        if use_rope:
            if rope_axes not in (2, 3):
                raise ValueError("rope_axes must be 2 or 3 when use_rope=True.")
            self._pair_count = hidden_size // (2 * rope_axes)
            self._rope_dim = self._pair_count * 2 * rope_axes
            if self._pair_count == 0:
                raise ValueError(f"hidden_size={hidden_size} too small for rope_axes={rope_axes}; need at least 2 * rope_axes channels.")
        else:
            self._pair_count, self._rope_dim = 0, 0
        if use_gate:
            self.gate_proj = nn.Linear(hidden_size, hidden_size, bias=True)
            self.gate_pos = nn.Parameter(torch.zeros(hidden_size))
        self.precision = "bf16"
        self._engine = None

    def _get_engine(self):
        from ..engine import EPDEngine
        if self._engine is None or not self._engine.is_bound():
            # the adapter must not register `self` as a child of itself -> keep it out of _modules
            object.__setattr__(self, "_stack", _ProcessorStack([self], self.hidden_size))
            self._engine = EPDEngine(self._stack)
        return self._engine

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor, size=None,
                pos: Optional[torch.Tensor] = None, phi: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        g = get_csr(edge_index, x.shape[0])
        if self.variant:
            from .. import variants
            if not x.is_cuda:
                raise RuntimeError("graphphysics_b200 runs on CUDA devices only (there is no CPU fallback)")
            x2, e2 = variants.graph_net_block_forward(self, x.float(), edge_attr.float()[g.perm_dst64], g, pos, phi, self.act,
                                                      3 if self.precision == "tight" else 1)
            return x2, e2[g.inv_perm_dst64]
        from ..engine import BlockFunction
        eng = self._get_engine()
        return BlockFunction.apply(eng.flat, x, edge_attr, eng, g)


# --------------------------------------------------------------------------------------------------
# Graph Transformer block (layers.py:213-278, 564-819 of the reference)
# --------------------------------------------------------------------------------------------------
class GatedMLP(nn.Module):
    """GELU(W1 x) * (W2 x)  (layers.py:213-249)."""

    def __init__(self, in_size: int, hidden_size: int, expansion_factor: int):
        super().__init__()
        self.linear1 = nn.Linear(in_size, expansion_factor * hidden_size)
        self.linear2 = nn.Linear(in_size, expansion_factor * hidden_size)
        self.act = "silu" if use_silu_activation() else "gelu"
        self.activation = nn.SiLU() if self.act == "silu" else nn.GELU()
        self.precision = "bf16"

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .. import dense
        if x.device.type != "cuda":
            raise RuntimeError("graphphysics_b200 runs on CUDA devices only (there is no CPU fallback)")
        return dense.gated_branch(x, None, None, self, None, add_resid=False, terms=3 if self.precision == "tight" else 1, act=self.act)


def build_gated_mlp(in_size: int, hidden_size: int, out_size: int, expansion_factor: int = 3) -> nn.Module:
    """RMSNorm -> GatedMLP -> Linear (layers.py:252-278); keys `0.scale`, `1.linear{1,2}.*`, `2.*`."""
    return nn.Sequential(RMSNorm(in_size), GatedMLP(in_size, hidden_size, expansion_factor),
                         nn.Linear(hidden_size * expansion_factor, out_size))


class Attention(nn.Module):
    """Multi-head attention masked by the mesh adjacency (layers.py:564-697).  q / k / v / proj run on the tensor
    cores (gp_gemm, bf16 operands, fp32 accumulate; q, k, v and y stored as bf16 in the reference's (N, d, heads)
    layout); scores, row softmax and the weighted sum -- the part the reference sends to dgl.sparse bsddmm / softmax /
    bspmm -- run in the CSR attention kernels of libgp_b200.so.  `adj` is the graph: a GraphCSR, or an edge_index
    tensor (2, E)."""

    def __init__(self, input_dim=512, output_dim=512, num_heads=4, pos_dimension: int = 3, use_proj_bias: bool = True,
                 use_separate_proj_weight: bool = True, use_rope_embeddings: bool = False,
                 use_gated_attention: bool = False, rope_base: float = 10000.0):
        super().__init__()
        assert output_dim % num_heads == 0, "Output dimension must be divisible by number of heads."
        self.hidden_size, self.num_heads, self.head_dim = output_dim, num_heads, output_dim // num_heads
        self.use_rope_embeddings, self.use_gated_attention = use_rope_embeddings, use_gated_attention
        self.pos_dimension, self.rope_base = pos_dimension, rope_base
        self.q_proj = nn.Linear(input_dim, output_dim, bias=use_proj_bias)
        self.k_proj = nn.Linear(input_dim, output_dim, bias=use_proj_bias)
        self.v_proj = nn.Linear(input_dim, output_dim, bias=use_proj_bias)
        self.proj = nn.Linear(output_dim, output_dim, bias=use_proj_bias)
        if use_rope_embeddings:
            self.m = self.head_dim // max(pos_dimension * 2, 1)
            step = math.log(rope_base) / max(self.m, 1)          # _make_inv_freq (layers.py:410-417)
            inv = torch.exp(-torch.arange(self.m, dtype=torch.float32) * step) if self.m > 0 else torch.empty(0)
            self.register_buffer("rope_inv_freq", inv, persistent=True)
        else:
            self.m = 0
            self.register_buffer("rope_inv_freq", torch.empty(0, dtype=torch.float32), persistent=False)
        self.gate_proj = nn.Linear(input_dim, output_dim, bias=use_proj_bias) if use_gated_attention else None
        self.precision = "bf16"        # "tight": three-term split GEMMs, fp32 q / k / v / y (set by the owning model)
        if not use_separate_proj_weight:
            with torch.no_grad():
                self.k_proj.weight = self.q_proj.weight
                self.v_proj.weight = self.q_proj.weight

    def _graph(self, x, adj):
        from ..graph import GraphCSR
        if adj is None:
            raise ValueError("Attention needs the mesh adjacency (a GraphCSR or an edge_index tensor)")
        return adj if isinstance(adj, GraphCSR) else get_csr(adj, x.shape[0])

    def forward(self, x: torch.Tensor, adj, pos: Optional[torch.Tensor] = None, return_attention: bool = False):
        from .. import dense
        g, terms = self._graph(x, adj), 3 if self.precision == "tight" else 1
        out = dense.attention_branch(x, None, self, g, add_resid=False, terms=terms, pos=self._rope_pos(pos))
        if return_attention:          # (out, attn) like layers.py:680-697; attn is computed beside the fused path
            return out, dense.attention_weights(self, x, None, g, terms=terms, pos=self._rope_pos(pos))
        return out

    def _rope_pos(self, pos):
        if not self.use_rope_embeddings:
            return None
        if pos is None:
            raise ValueError("RoPE embeddings require positional information when enabled.")
        return pos


class Transformer(nn.Module):
    """x += Attention(norm1(x));  x += gated_mlp(norm2(x))   (layers.py:700-819)."""

    def __init__(self, input_dim: int, output_dim: int, num_heads: int, activation_layer=nn.ReLU, use_proj_bias: bool = True,
                 use_separate_proj_weight: bool = True, use_rope_embeddings: bool = False,
                 use_gated_attention: bool = False, pos_dimension: int = 3, rope_base: float = 10000.0):
        super().__init__()
        self.use_rope_embeddings, self.use_gated_attention, self.pos_dimension = (use_rope_embeddings, use_gated_attention,
                                                                                pos_dimension)
        self.attention = Attention(input_dim=input_dim, output_dim=output_dim, num_heads=num_heads,
                                   pos_dimension=pos_dimension, use_proj_bias=use_proj_bias,
                                   use_separate_proj_weight=use_separate_proj_weight,
                                   use_rope_embeddings=use_rope_embeddings, use_gated_attention=use_gated_attention,
                                   rope_base=rope_base)
        self.activation = activation_layer()        # constructed but unused, as in the reference (layers.py:757)
        self.norm1, self.norm2 = RMSNorm(output_dim), RMSNorm(output_dim)
        self.gated_mlp = build_gated_mlp(in_size=output_dim, hidden_size=output_dim, out_size=output_dim)
        self.use_adjacency = True
        self.set_precision("bf16")

    def set_precision(self, precision: str) -> None:
        """"bf16": bf16 MMA operands / bf16 q, k, v, y; "tight": three-term split GEMMs with fp32 tensors (csrc/gemm.cu)."""
        if precision not in ("bf16", "tight"):
            raise ValueError(f"precision must be 'bf16' or 'tight', got {precision!r}")
        self.precision = self.attention.precision = self.gated_mlp[1].precision = precision

    def forward(self, x: torch.Tensor, adj, pos: Optional[torch.Tensor] = None, return_attention: bool = False):
        """x (fp32 residual stream) -> x + Attn(norm1(x)) -> ... + W3(GELU(W1 n) * (W2 n)), n = norm_g(norm2(.)): seven
        GEMMs, the CSR attention kernel and three row-wise kernels; both residual adds ride in a GEMM epilogue.  Two
        autograd nodes (dense.attention_branch, dense.gated_branch) with hand-written backwards."""
        from .. import dense
        terms = 3 if self.precision == "tight" else 1
        attn = None
        if return_attention:          # layers.py:795-801: the attention matrix of this block's input
            attn = dense.attention_weights(self.attention, x, self.norm1.scale, self.attention._graph(x, adj), terms=terms,
                                           pos=self.attention._rope_pos(pos))
        x = dense.attention_branch(x, self.norm1.scale, self.attention, self.attention._graph(x, adj), add_resid=True, terms=terms,
                                   pos=self.attention._rope_pos(pos))
        # the double norm: Transformer.norm2, then build_gated_mlp's own leading RMSNorm (layers.py:252-278)
        x = dense.gated_branch(x, self.norm2.scale, self.gated_mlp[0].scale, self.gated_mlp[1], self.gated_mlp[2],
                               add_resid=True, terms=terms, act=self.gated_mlp[1].act)
        return (x, attn) if return_attention else x


class TemporalAttention(nn.Module):
    """Temporal corrector (graphphysics/models/layers.py:822-887): adjacency-masked cross-attention with queries and
    values from the last block's output `h_pred` and keys from its input `h_prev`, a sigmoid gate on [h_pred | h_prev],
    the residual onto h_prev and a two-layer SiLU mixer on [h_corr | h_prev].  Same constructor and state_dict keys
    (q_proj / k_proj / v_proj / out_proj / gate.{0,2} / mixer.{0,2}).  Every Linear is gp_gemm, the attention the CSR
    kernels of csrc/attention.cu, SiLU / sigmoid-multiply / concatenation the row kernels of csrc/variant_ops.cu
    (composed in graphphysics_b200/variants.py::temporal_attention_forward)."""

    def __init__(self, hidden_size: int, num_heads: int = 4, use_gate: bool = True):
        super().__init__()
        assert hidden_size % num_heads == 0, "hidden_size must be divisible by num_heads"
        self.h, self.H, self.d = hidden_size, num_heads, hidden_size // num_heads
        self.use_gate = use_gate
        self.q_proj = nn.Linear(self.h, self.h, bias=True)
        self.k_proj = nn.Linear(self.h, self.h, bias=True)
        self.v_proj = nn.Linear(self.h, self.h, bias=True)
        self.out_proj = nn.Linear(self.h, self.h, bias=True)
        if use_gate:
            self.gate = nn.Sequential(nn.Linear(2 * self.h, self.h), nn.SiLU(), nn.Linear(self.h, self.h), nn.Sigmoid())
        self.mixer = nn.Sequential(nn.Linear(2 * self.h, self.h), nn.SiLU(), nn.Linear(self.h, self.h))
        self.precision = "bf16"

    def forward(self, h_prev: torch.Tensor, h_pred: torch.Tensor, adj=None) -> torch.Tensor:
        """adj: a GraphCSR, or an edge_index [2, E] (rows = queries edge_index[0], columns = keys edge_index[1], as
        dglsp.spmatrix(indices=edge_index) at processors.py:183 / 366)."""
        from ..graph import GraphCSR
        from ..variants import temporal_attention_forward
        if adj is None:
            raise ValueError("TemporalAttention needs the adjacency (the reference's DGL sparse matrix)")
        if not h_prev.is_cuda:
            raise RuntimeError("graphphysics_b200 runs on CUDA devices only (there is no CPU fallback)")
        g = adj if isinstance(adj, GraphCSR) else get_csr(adj, h_prev.shape[0])
        return temporal_attention_forward(self, h_prev, h_pred, g, 3 if self.precision == "tight" else 1)
