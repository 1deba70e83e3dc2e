"""Variant flags of the path on native kernels: use_silu_activation, use_gated_mlp, use_gate (the
query-conditioned gate on the aggregated messages) and use_rope (relative rotary embedding of the senders) for
GraphNetBlock / EncodeProcessDecode (graphphysics/models/layers.py:890-1149, processors.py:57-215).

No shipped training_config switches these on, so they do not get fused kernels of their own: a block is composed from
the tensor-core GEMM (gp_gemm), the row gather / segment-sum kernels and the row-wise kernels of csrc/variant_ops.cu,
each wrapped in a torch.autograd.Function with a hand-written backward.  All tensors between kernels are fp32; GEMM
operands (activations, weights, incoming gradients) are rounded to bf16 inside gp_gemm, fp32 accumulation -- the
arithmetic oracle/gp_oracle.py::epd_forward_variant(mode="bf16") restates.  precision="tight" uses the three-term split
(terms = 3) instead.  Edges are processed in receiver-sorted order (GraphCSR) so the scatter-sum is a contiguous,
atomic-free segment sum and the gathers' backward a segment sum as well.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import dense, ops
from ._lib import check, lib, ptr, stream_ptr

ACT_KIND = {"relu": 1, "silu": 2, "gelu": 3}


def _c(t):
    return C.c_void_p(ptr(t))


def _st():
    return C.c_void_p(stream_ptr())


# ------------------------------------------------------------------------------------------------ small functions
class _Linear(torch.autograd.Function):
    """y = x W^T + b, fp32 in / out, operands rounded to bf16 inside gp_gemm (or split in three terms)."""

    @staticmethod
    def forward(ctx, x, w, b, terms: int):
        x = dense._f32(x)
        ctx.save_for_backward(x, w)
        ctx.has_bias, ctx.terms = b is not None, terms
        return dense.lin_fwd(x, w, b, terms=terms)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dense._f32(dy)
        dx = dense.lin_dgrad(dy, w, terms=ctx.terms) if ctx.needs_input_grad[0] else None
        if ctx.has_bias:
            dw, db = dense.lin_wgrad(dy, x, bias=True, terms=ctx.terms)
        else:
            dw, db = dense.lin_wgrad(dy, x, terms=ctx.terms), None
        return dx, dw, db, None


def linear(x, lin, terms: int = 1):
    return _Linear.apply(x, lin.weight, lin.bias, terms)


class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, kind: int):
        z = dense._f32(z)
        out = torch.empty_like(z)
        check(lib().gp_act_fwd(_c(z), C.c_int64(z.numel()), C.c_int32(kind), C.c_void_p(None), _c(out), _st()), "gp_act_fwd")
        ops._launched()
        ctx.save_for_backward(z)
        ctx.kind = kind
        return out

    @staticmethod
    def backward(ctx, d):
        (z,) = ctx.saved_tensors
        d = dense._f32(d).clone()
        check(lib().gp_act_bwd(_c(z), C.c_int64(z.numel()), C.c_int32(ctx.kind), _c(d), _st()), "gp_act_bwd")
        ops._launched()
        return d, None


class _Glu(torch.autograd.Function):
    """act(a1) * a2 (GatedMLP, layers.py:213-249)."""

    @staticmethod
    def forward(ctx, a1, a2, kind: int):
        a1, a2 = dense._f32(a1), dense._f32(a2)
        R, G = a1.shape
        out = torch.empty_like(a1)
        check(lib().gp_glu_fwd(_c(a1), _c(a2), C.c_int32(G), C.c_int64(R), C.c_int32(G), C.c_int32(kind), C.c_void_p(None), _c(out), _st()),
              "gp_glu_fwd")
        ops._launched()
        ctx.save_for_backward(a1, a2)
        ctx.kind = kind
        return out

    @staticmethod
    def backward(ctx, dg):
        a1, a2 = ctx.saved_tensors
        dg = dense._f32(dg)
        R, G = a1.shape
        da1, da2 = torch.empty_like(a1), torch.empty_like(a2)
        check(lib().gp_glu_bwd(_c(a1), _c(a2), C.c_int32(G), _c(dg), C.c_int64(R), C.c_int32(G), C.c_int32(ctx.kind), _c(da1), _c(da2),
                               C.c_int32(G), _st()), "gp_glu_bwd")
        ops._launched()
        return da1, da2, None


def _segsum(src: torch.Tensor, perm: Optional[torch.Tensor], rowptr: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.empty((n, src.shape[1]), dtype=torch.float32, device=src.device)
    check(lib().gp_segsum_rows_f32(_c(src), C.c_int32(src.stride(0)), _c(perm), _c(rowptr), C.c_int64(n), C.c_int32(src.shape[1]), _c(out), _st()),
          "gp_segsum_rows_f32")
    ops._launched()
    return out


def _gather(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    out = torch.empty((idx.numel(), x.shape[1]), dtype=x.dtype, device=x.device)
    ops.halo_pack(x, idx, out)
    return out


class _GatherRows(torch.autograd.Function):
    """out[e] = x[idx[e]]; backward: dx[n] = sum of dout over the entries with idx == n, walked through (perm, rowptr)."""

    @staticmethod
    def forward(ctx, x, idx, perm, rowptr):
        x = dense._f32(x)
        ctx.save_for_backward(perm, rowptr)
        ctx.n = x.shape[0]
        return _gather(x, idx)

    @staticmethod
    def backward(ctx, d):
        perm, rowptr = ctx.saved_tensors
        return _segsum(dense._f32(d), perm, rowptr, ctx.n), None, None, None


class _SegmentSum(torch.autograd.Function):
    """agg[n] = sum of msg over the receiver-sorted segment of n (PyG propagate aggr="add", layers.py:926, 1031-1037)."""

    @staticmethod
    def forward(ctx, msg, rowptr, dst):
        msg = dense._f32(msg)
        ctx.save_for_backward(dst)
        return _segsum(msg, None, rowptr, rowptr.numel() - 1)

    @staticmethod
    def backward(ctx, d):
        (dst,) = ctx.saved_tensors
        return _gather(dense._f32(d), dst), None, None


class _RopeRel(torch.autograd.Function):
    """Relative rotary embedding of the gathered sender rows (layers.py:1104-1149); pos is not differentiated (the
    reference's positions are data)."""

    @staticmethod
    def forward(ctx, x_src, pos, src, dst, axes: int, pair_count: int, base: float):
        x_src = dense._f32(x_src)
        ctx.save_for_backward(pos, src, dst)
        ctx.cfg = (axes, pair_count, base)
        return _RopeRel._run(x_src, pos, src, dst, axes, pair_count, base, 0)

    @staticmethod
    def _run(x, pos, src, dst, axes, pc, base, inverse):
        out = torch.empty_like(x)
        check(lib().gp_rope_rel(_c(x), _c(pos), C.c_int32(pos.stride(0)), _c(src), _c(dst), C.c_int64(x.shape[0]), C.c_int32(x.shape[1]),
                                C.c_int32(axes), C.c_int32(pc), C.c_float(base), C.c_int32(inverse), _c(out), _st()), "gp_rope_rel")
        ops._launched()
        return out

    @staticmethod
    def backward(ctx, d):
        pos, src, dst = ctx.saved_tensors
        axes, pc, base = ctx.cfg
        return _RopeRel._run(dense._f32(d), pos, src, dst, axes, pc, base, 1), None, None, None, None, None, None


class _SigmoidMul(torch.autograd.Function):
    """v * sigmoid(logits) (the aggregation gate, layers.py:1091-1098; the attention gate, layers.py:684-689)."""

    @staticmethod
    def forward(ctx, logits, v):
        logits, v = dense._f32(logits), dense._f32(v)
        out = torch.empty_like(v)
        check(lib().gp_sigmoid_mul_fwd(_c(logits), _c(v), C.c_int64(v.numel()), _c(out), _st()), "gp_sigmoid_mul_fwd")
        ops._launched()
        ctx.save_for_backward(logits, v)
        return out

    @staticmethod
    def backward(ctx, d):
        logits, v = ctx.saved_tensors
        d = dense._f32(d)
        dl, dv = torch.empty_like(logits), torch.empty_like(v)
        check(lib().gp_sigmoid_mul_bwd(_c(logits), _c(v), _c(d), C.c_int64(v.numel()), _c(dl), _c(dv), _st()), "gp_sigmoid_mul_bwd")
        ops._launched()
        return dl, dv


class _AddOuter(torch.autograd.Function):
    """logits + phi[:, None] * vec[None, :] (gate_pos term of the aggregation gate); phi is data."""

    @staticmethod
    def forward(ctx, logits, phi, vec):
        out = dense._f32(logits).clone()
        phi = dense._f32(phi.reshape(-1))
        check(lib().gp_add_outer(_c(out), _c(phi), _c(vec), C.c_int64(out.shape[0]), C.c_int32(out.shape[1]), _st()), "gp_add_outer")
        ops._launched()
        ctx.save_for_backward(phi)
        return out

    @staticmethod
    def backward(ctx, d):
        (phi,) = ctx.saved_tensors
        d = dense._f32(d)
        # dvec[c] = sum_r d[r, c] * phi[r]: the bias-gradient product with phi in place of the ones column
        dvec = dense.lin_wgrad(d, phi.reshape(-1, 1).contiguous(), terms=3).reshape(-1)
        return d, None, dvec


class _Concat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        a, b = dense._f32(a), dense._f32(b)
        c = dense._f32(c) if c is not None else None
        wa, wb, wc = a.shape[1], b.shape[1], (c.shape[1] if c is not None else 0)
        out = torch.empty((a.shape[0], wa + wb + wc), dtype=torch.float32, device=a.device)
        check(lib().gp_concat_rows(_c(a), C.c_int32(wa), _c(b), C.c_int32(wb), _c(c), C.c_int32(wc), C.c_int64(a.shape[0]), _c(out), _st()),
              "gp_concat_rows")
        ops._launched()
        ctx.widths = (wa, wb, wc)
        return out

    @staticmethod
    def backward(ctx, d):
        d = dense._f32(d)
        outs, col = [], 0
        for w in ctx.widths:
            if w == 0:
                outs.append(None)
                continue
            o = torch.empty((d.shape[0], w), dtype=torch.float32, device=d.device)
            check(lib().gp_split_cols(_c(d), C.c_int32(d.shape[1]), C.c_int32(col), C.c_int32(w), C.c_int64(d.shape[0]), _c(o), C.c_int32(0), _st()),
                  "gp_split_cols")
            ops._launched()
            outs.append(o)
            col += w
        return tuple(outs)


# ------------------------------------------------------------------------------------------------ modules' forward passes
def mlp_seq(seq, x, act: str, terms: int = 1):
    """build_mlp container with the activation `act` (layers.py:163-210)."""
    kind = ACT_KIND[act]
    h = x
    n_lin = sum(1 for m in seq if isinstance(m, torch.nn.Linear))
    li = 0
    for m in seq:
        if isinstance(m, torch.nn.Linear):
            h = linear(h, m, terms)
            li += 1
            if li < n_lin:
                h = _Act.apply(h, kind)
        elif hasattr(m, "scale"):
            h = dense.rms_norm(h, m.scale)
    return h


def gated_mlp_seq(seq, x, act: str, terms: int = 1):
    """build_gated_mlp container: RMSNorm -> GatedMLP -> Linear (layers.py:252-278); SiLU gating under the global flag."""
    n = dense.rms_norm(x, seq[0].scale)
    a1, a2 = linear(n, seq[1].linear1, terms), linear(n, seq[1].linear2, terms)
    return linear(_Glu.apply(a1, a2, ACT_KIND["silu" if act == "silu" else "gelu"]), seq[2], terms)


def graph_net_block_forward(block, x, e_sorted, g, pos, phi, act: str, terms: int = 1):
    """GraphNetBlock.forward (layers.py:989-1042) on receiver-sorted edges: returns (x', e'_sorted)."""
    x_i = _GatherRows.apply(x, g.dst, None, g.rowptr_dst)                     # receivers: contiguous segments
    x_j = _GatherRows.apply(x, g.src, g.perm_src, g.rowptr_src)
    if block.use_rope:
        if pos is None:
            raise ValueError("Node positions `pos` must be provided when use_rope=True.")
        p = pos[:, :block.rope_axes].float().contiguous()
        x_j = _RopeRel.apply(x_j, p, g.src, g.dst, block.rope_axes, block._pair_count, float(block.rope_base))
    cat = _Concat.apply(e_sorted, x_i, x_j)
    run = gated_mlp_seq if block.use_gated_mlp else mlp_seq
    e_upd = run(block.edge_block, cat, act, terms)
    agg = _SegmentSum.apply(e_upd, g.rowptr_dst, g.dst)
    if block.use_gate:
        logits = linear(x, block.gate_proj, terms)
        if phi is not None:
            logits = _AddOuter.apply(logits, phi.to(x.device), block.gate_pos)
        agg = _SigmoidMul.apply(logits, agg)
    x_upd = run(block.node_block, _Concat.apply(x, agg, None), act, terms)
    return x + x_upd, e_sorted + e_upd


def temporal_attention_forward(blk, h_prev, h_pred, g, terms: int = 1):
    """TemporalAttention.forward (layers.py:861-887).  q, v from h_pred, k from h_prev; (N, H) -> (N, head_dim, heads)
    is the channel layout the attention kernels use (c = d * heads + h)."""
    h_prev, h_pred = dense._f32(h_prev), dense._f32(h_pred)
    q, k, v = linear(h_pred, blk.q_proj, terms), linear(h_prev, blk.k_proj, terms), linear(h_pred, blk.v_proj, terms)
    out = linear(ops.CSRAttention.apply(q, k, v, g, blk.H), blk.out_proj, terms)
    if blk.use_gate:
        z = _Act.apply(linear(_Concat.apply(h_pred, h_prev, None), blk.gate[0], terms), ACT_KIND["silu"])
        out = _SigmoidMul.apply(linear(z, blk.gate[2], terms), out)
    h_corr = h_prev + out
    z = _Act.apply(linear(_Concat.apply(h_corr, h_prev, None), blk.mixer[0], terms), ACT_KIND["silu"])
    return h_corr + linear(z, blk.mixer[2], terms)


def epd_forward(model, graph, act: str):
    """EncodeProcessDecode.forward with variant flags (processors.py:162-215)."""
    from .graph import get_csr
    x, edge_attr = graph.x, graph.edge_attr
    if not x.is_cuda:
        raise RuntimeError("graphphysics_b200 runs on CUDA devices only (there is no CPU fallback)")
    terms = 3 if model.precision == "tight" else 1
    g = get_csr(graph.edge_index, x.shape[0])
    e = edge_attr.float()[g.perm_dst64]
    x = x.float()
    if not model.only_processor:
        x = mlp_seq(model.nodes_encoder, x, act, terms)
        e = mlp_seq(model.edges_encoder, e, act, terms)
    pos = getattr(graph, "pos", None) if model.use_rope else None
    if model.use_rope and pos is None:
        raise ValueError("Graph data must contain `pos` when use_rope_embeddings=True.")
    phi = getattr(graph, "phi", None) if model.use_gate else None
    prev_x = x
    for blk in model.processor_list:
        prev_x = x
        x, e = graph_net_block_forward(blk, x, e, g, pos, phi, act, terms)
    if model.temporal_block is not None:                    # processors.py:204-209: (input, output) of the last block
        x = temporal_attention_forward(model.temporal_block, prev_x, x, g, terms)
    if model.only_processor:
        return x
    return mlp_seq(model.decode_module, x, act, terms)
