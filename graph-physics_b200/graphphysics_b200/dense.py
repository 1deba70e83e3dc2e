"""Dense side of the graph-Transformer path on the native kernels (no library GEMM, no eager elementwise math):

    Linear      y = x W^T + b [+ resid] [relu]          gp_gemm (tcgen05, bf16 operands, fp32 accumulate)
    RMSNorm     one norm, or norm2 followed by the gated MLP's own norm           gp_rmsnorm_fwd / _bwd
    GeluGate    GELU(a1) * a2                                                     gp_gelu_gate_fwd / _bwd
    bias / scale gradients: per-block partial sums + fixed-order reduction        gp_colsum, gp_reduce_partials

Three torch.autograd.Functions with hand-written backwards cover the model (graphphysics/models/layers.py:104-129,
163-278, 637-697, 766-819):

    attention_branch   [x +] proj(attention(q, k, v of [norm1](x)))       Attention.forward / first half of Transformer.forward
    gated_branch       [x +] [W3] (GELU(W1 n) * (W2 n)), n = [norms](x)   GatedMLP.forward / second half of Transformer.forward
    mlp4               build_mlp: Linear, ReLU x3, Linear, [RMSNorm]      nodes_encoder / decode_module

Every bf16 tensor (norm outputs, q / k / v / y, the gate output, hidden activations) lives INSIDE one of these functions,
so torch never sees a bf16 tensor that needs a gradient (it would cast that gradient to bf16): all gradients between
kernels are fp32, rounded to bf16 only as MMA operands.  Arithmetic ("kernel specification", restated by
oracle/gp_oracle.py mode="bf16"): GEMM operands -- activations, weights and, in the backward, the incoming gradient --
rounded to bf16, fp32 accumulation; the residual stream x, norm statistics, GELU and all parameter-gradient sums in
fp32.  terms = 3 (precision="tight") splits every fp32 operand into three bf16 terms instead and keeps all tensors fp32.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import ops
from ._lib import AttentionArgs, GemmArgs, check, lib, ptr, stream_ptr


# ------------------------------------------------------------------------------------------------ kernel wrappers (no autograd)
def gemm(M: int, N: int, K: int, a: torch.Tensor, a_sm: int, a_sk: int, b: torch.Tensor, b_sn: int, b_sk: int, c: torch.Tensor,
         c_sm: int, c_sn: int, *, bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None, relu: bool = False,
         accumulate: bool = False, split_k: int = 1, terms: int = 1, b_ones: bool = False) -> None:
    """C(m,n) = [C +] [resid +] bias[n] + sum_k A(m,k) B(n,k) with strided fp32 / bf16 operands (gp_gemm).  b_ones: row N-1
    of B is not in memory and reads as 1.0."""
    for t in (a, b, c):
        assert t.is_cuda and t.dtype in (torch.float32, torch.bfloat16)
    g = GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.a, g.a_sm, g.a_sk = ptr(a), a_sm, a_sk
    g.b, g.b_sn, g.b_sk = ptr(b), b_sn, b_sk
    g.c, g.c_sm, g.c_sn = ptr(c), c_sm, c_sn
    g.bias, g.resid = ptr(bias), ptr(resid)
    g.a_bf16, g.b_bf16, g.c_bf16 = (int(t.dtype == torch.bfloat16) for t in (a, b, c))
    g.relu, g.accumulate, g.terms, g.split_k, g.b_ones = int(relu), int(accumulate), terms, split_k, int(b_ones)
    part = None
    if split_k > 1:
        part = torch.empty(split_k * M * ((N + 3) // 4 * 4), dtype=torch.float32, device=c.device)
        g.partials = ptr(part)
    ev = ops.PROFILE.begin("gemm")
    check(lib().gp_gemm(C.byref(g), C.c_void_p(stream_ptr())), "gp_gemm")
    ops.PROFILE.end("gemm", ev)
    ops._launched(2 if split_k > 1 else 1)


def _reduce_rows(part: torch.Tensor, n_parts: int, cols: int) -> torch.Tensor:
    out = torch.empty(cols, dtype=torch.float32, device=part.device)
    ops.reduce_partials(part, n_parts, cols, 0, 1, cols, cols, out, cols, False)
    return out


def colsum(src: torch.Tensor, round_bf16: bool) -> torch.Tensor:
    """Column sums of a contiguous fp32 [rows, cols] matrix (bias gradient), fixed order."""
    rows, cols = src.shape
    nb = int(lib().gp_colsum_blocks(C.c_int32(rows)))
    part = torch.empty((nb, cols), dtype=torch.float32, device=src.device)
    check(lib().gp_colsum(C.c_void_p(ptr(src)), C.c_int32(src.stride(0)), C.c_int32(rows), C.c_int32(cols), C.c_int32(int(round_bf16)),
                          C.c_void_p(ptr(part)), C.c_void_p(stream_ptr())), "gp_colsum")
    ops._launched()
    return _reduce_rows(part, nb, cols)


def _split_k(rows: int, tiles: int = 1) -> int:
    """CTAs along the contraction of a wgrad: up to four co-resident CTAs per SM (one 64-row chunk per CTA at least), so
    the load -> MMA -> store chain of one CTA hides behind the others'."""
    chunks = (rows + 63) // 64
    return max(1, min(chunks, 592 // max(tiles, 1), 1024))


def lin_fwd(x, w, b, *, resid=None, relu=False, out_bf16=False, terms=1) -> torch.Tensor:
    """y = x W^T + b [+ resid] [relu]; x [R, K] fp32 or bf16 (rows contiguous), w [N, K]."""
    R, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and w.stride(1) == 1 and x.stride(1) == 1
    y = torch.empty((R, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
    gemm(R, N, K, x, x.stride(0), 1, w, w.stride(0), 1, y, N, 1, bias=b, resid=resid, relu=relu, terms=terms)
    return y


def lin_dgrad(dy, w, *, out=None, terms=1) -> torch.Tensor:
    """dx[r, k] (+)= sum_n dy[r, n] w[n, k]; accumulates into `out` when given."""
    R, N = dy.shape
    K = w.shape[1]
    acc = out is not None
    if out is None:
        out = torch.empty((R, K), dtype=torch.float32, device=dy.device)
    gemm(R, K, N, dy, N, 1, w, 1, w.stride(0), out, K, 1, accumulate=acc, terms=terms)
    return out


def lin_wgrad(dy, x, *, bias: bool = False, terms=1):
    """dw[n, k] = sum_r dy[r, n] x[r, k] (split over CTAs along r, reduced in fixed order).  bias=True: x is read with one
    more column of ones, so column K of the same product is db[n] = sum_r dy[r, n] -- returns (dw, db), views of one buffer."""
    R, N = dy.shape
    K = x.shape[1]
    if not bias:
        dw = torch.empty((N, K), dtype=torch.float32, device=dy.device)
        tiles = ((N + 127) // 128) * ((K + 127) // 128)
        gemm(N, K, R, dy, 1, N, x, 1, x.stride(0), dw, K, 1, split_k=_split_k(R, tiles), terms=terms)
        return dw
    ld = K + 8                                                   # padded row: 16-byte aligned rows for the vector stores
    buf = torch.empty((N, ld), dtype=torch.float32, device=dy.device)
    tiles = ((N + 127) // 128) * ((K + 1 + 127) // 128)
    gemm(N, K + 1, R, dy, 1, N, x, 1, x.stride(0), buf, ld, 1, split_k=_split_k(R, tiles), terms=terms, b_ones=True)
    return buf[:, :K], buf[:, K]


def bias_grad(dy, terms=1) -> torch.Tensor:
    return colsum(dy, round_bf16=(terms == 1))


def relu_mask_(d: torch.Tensor, h: torch.Tensor) -> None:
    """d[i] = h[i] > 0 ? d[i] : 0 in place (d is a buffer this module allocated)."""
    check(lib().gp_relu_bwd(C.c_void_p(ptr(d)), C.c_void_p(ptr(h)), C.c_int32(int(h.dtype == torch.bfloat16)), C.c_int64(d.numel()),
                            C.c_void_p(stream_ptr())), "gp_relu_bwd")
    ops._launched()


def norm_fwd(x, s1, s2=None, out_bf16=True) -> torch.Tensor:
    R, H = x.shape
    out = torch.empty((R, H), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
    check(lib().gp_rmsnorm_fwd(C.c_void_p(ptr(x)), C.c_int32(x.stride(0)), C.c_int32(R), C.c_int32(H), C.c_void_p(ptr(s1)),
                               C.c_void_p(ptr(s2)), C.c_void_p(ptr(out) if out_bf16 else None), C.c_void_p(None if out_bf16 else ptr(out)),
                               C.c_int32(H), C.c_void_p(stream_ptr())), "gp_rmsnorm_fwd")
    ops._launched()
    return out


def norm_bwd(x, s1, s2, dy, add=None):
    """dx = [add +] J^T dy through norm(x; s1) [then norm(.; s2)]; returns dx, dscale1, dscale2 (None without s2)."""
    R, H = x.shape
    nb = int(lib().gp_rmsnorm_bwd_blocks(C.c_int32(R)))
    p1 = torch.empty((nb, H), dtype=torch.float32, device=x.device)
    p2 = torch.empty((nb, H), dtype=torch.float32, device=x.device) if s2 is not None else None
    dx = torch.empty((R, H), dtype=torch.float32, device=x.device)
    check(lib().gp_rmsnorm_bwd(C.c_void_p(ptr(x)), C.c_int32(x.stride(0)), C.c_int32(R), C.c_int32(H), C.c_void_p(ptr(s1)),
                               C.c_void_p(ptr(s2)), C.c_void_p(ptr(dy)), C.c_int32(dy.stride(0)), C.c_void_p(ptr(dx)), C.c_int32(H),
                               C.c_void_p(ptr(add)), C.c_int32(add.stride(0) if add is not None else 0), C.c_void_p(ptr(p1)),
                               C.c_void_p(ptr(p2)), C.c_void_p(stream_ptr())), "gp_rmsnorm_bwd")
    ops._launched()
    return dx, _reduce_rows(p1, nb, H), (_reduce_rows(p2, nb, H) if s2 is not None else None)


def gelu_fwd(a1, a2, out_bf16=True, kind: int = 3) -> torch.Tensor:
    """act(a1) * a2; kind 3 = GELU (gp_gelu_gate_fwd), 2 = SiLU (gp_glu_fwd)."""
    g = torch.empty(a1.shape, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=a1.device)
    if kind != 3 or not a1.is_contiguous():      # SiLU gating, or a1 | a2 as column blocks of one buffer (row stride ld)
        R, G = a1.shape
        assert a1.stride() == a2.stride() and a1.stride(1) == 1
        check(lib().gp_glu_fwd(C.c_void_p(ptr(a1)), C.c_void_p(ptr(a2)), C.c_int32(a1.stride(0)), C.c_int64(R), C.c_int32(G), C.c_int32(kind),
                               C.c_void_p(ptr(g) if out_bf16 else None), C.c_void_p(None if out_bf16 else ptr(g)), C.c_void_p(stream_ptr())),
              "gp_glu_fwd")
        ops._launched()
        return g
    check(lib().gp_gelu_gate_fwd(C.c_void_p(ptr(a1)), C.c_void_p(ptr(a2)), C.c_int64(a1.numel()), C.c_void_p(ptr(g) if out_bf16 else None),
                                 C.c_void_p(None if out_bf16 else ptr(g)), C.c_void_p(stream_ptr())), "gp_gelu_gate_fwd")
    ops._launched()
    return g


def gelu_bwd(a1, a2, dg, kind: int = 3, fused_out: bool = False):
    """fused_out: da1 | da2 are the column blocks of ONE [R, 2G] buffer (returned as two views)."""
    R, G = a1.shape
    if fused_out:
        da12 = torch.empty((R, 2 * G), dtype=torch.float32, device=a1.device)
        da1, da2 = da12[:, :G], da12[:, G:]
    else:
        da1, da2 = torch.empty((R, G), dtype=torch.float32, device=a1.device), torch.empty((R, G), dtype=torch.float32, device=a1.device)
    if kind != 3 or fused_out or not a1.is_contiguous():
        check(lib().gp_glu_bwd(C.c_void_p(ptr(a1)), C.c_void_p(ptr(a2)), C.c_int32(a1.stride(0)), C.c_void_p(ptr(dg)), C.c_int64(R), C.c_int32(G),
                               C.c_int32(kind), C.c_void_p(ptr(da1)), C.c_void_p(ptr(da2)), C.c_int32(da1.stride(0)), C.c_void_p(stream_ptr())),
              "gp_glu_bwd")
        ops._launched()
        return da1, da2
    check(lib().gp_gelu_gate_bwd(C.c_void_p(ptr(a1)), C.c_void_p(ptr(a2)), C.c_void_p(ptr(dg)), C.c_int64(a1.numel()), C.c_void_p(ptr(da1)),
                                 C.c_void_p(ptr(da2)), C.c_void_p(stream_ptr())), "gp_gelu_gate_bwd")
    ops._launched()
    return da1, da2


def attn_fwd(q, k, v, g, num_heads: int):
    """Adjacency-masked multi-head attention over the CSR rows of `g`; q, k, v [N, H] all fp32 or all bf16."""
    n, h = q.shape
    a = AttentionArgs()
    a.io_bf16 = int(q.dtype == torch.bfloat16)
    a.n, a.hidden, a.num_heads = n, h, num_heads
    a.q, a.k, a.v = ptr(q), ptr(k), ptr(v)
    assert q.stride(1) == 1 and q.stride() == k.stride() == v.stride()
    a.ld_qkv = q.stride(0)                      # h, or 3h when q | k | v are column blocks of one buffer
    a.rowptr, a.col = ptr(g.rowptr_src), ptr(g.att_col)
    y = torch.empty((n, h), dtype=q.dtype, device=q.device)
    y32 = torch.empty(q.shape, dtype=torch.float32, device=q.device) if a.io_bf16 else None
    lse = torch.empty((n, num_heads), dtype=torch.float32, device=q.device)
    a.y, a.lse, a.y_f32 = ptr(y), ptr(lse), ptr(y32)
    ev = ops.PROFILE.begin("attn_fwd")
    check(lib().gp_csr_attention_fwd(C.byref(a), C.c_void_p(stream_ptr())), "gp_csr_attention_fwd")
    ops.PROFILE.end("attn_fwd", ev)
    ops._launched()
    return y, y32, lse


def attn_bwd(q, k, v, y, y32, lse, dy, g, num_heads: int, fused_out: bool = False):
    """fused_out: dq | dk | dv are written as the column blocks of ONE [N, 3h] buffer (returned as three views)."""
    n, h = q.shape
    a = AttentionArgs()
    a.io_bf16 = int(q.dtype == torch.bfloat16)
    a.n, a.hidden, a.num_heads = n, h, num_heads
    a.ld_qkv = q.stride(0)
    a.q, a.k, a.v, a.y, a.lse, a.dy, a.y_f32 = ptr(q), ptr(k), ptr(v), ptr(y), ptr(lse), ptr(dy), ptr(y32)
    a.rowptr, a.col, a.pos = ptr(g.rowptr_src), ptr(g.att_col), ptr(g.perm_src)
    a.colptr, a.row = ptr(g.rowptr_dst), ptr(g.src)
    if fused_out:
        dqkv = torch.empty((n, 3 * h), dtype=torch.float32, device=q.device)
        dq, dk, dv = dqkv[:, :h], dqkv[:, h:2 * h], dqkv[:, 2 * h:]
        a.ld_dqkv = 3 * h
    else:
        dq, dk, dv = (torch.empty((n, h), dtype=torch.float32, device=q.device) for _ in range(3))
    ea = torch.empty((g.num_edges, num_heads), dtype=torch.float32, device=q.device)
    eds = torch.empty_like(ea)
    a.dq, a.dk, a.dv, a.edge_a, a.edge_ds = ptr(dq), ptr(dk), ptr(dv), ptr(ea), ptr(eds)
    ev = ops.PROFILE.begin("attn_bwd")
    check(lib().gp_csr_attention_bwd(C.byref(a), C.c_void_p(stream_ptr())), "gp_csr_attention_bwd")
    ops.PROFILE.end("attn_bwd", ev)
    ops._launched(2)
    return dq, dk, dv


def _stackable(ws, bs) -> bool:
    """The weights (and, if present, the biases) are consecutive slices of one buffer -- engine.FlatParams lays q / k / v
    and linear1 / linear2 out that way -- so they read as one stacked matrix / vector."""
    from .engine import _adjacent
    if any(w.shape != ws[0].shape or not w.is_contiguous() for w in ws) or not _adjacent(*ws):
        return False
    if all(b is None for b in bs):
        return True
    return all(b is not None for b in bs) and _adjacent(*bs)


def _stack(first: torch.Tensor, k: int) -> torch.Tensor:
    """View of `k` adjacent equal-shaped tensors starting at `first` as one tensor stacked along dim 0."""
    shape = (k * first.shape[0],) + tuple(first.shape[1:])
    return torch.as_strided(first.detach(), shape, first.stride())


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.device.type != "cuda":
        raise RuntimeError("graphphysics_b200 runs on CUDA devices only (there is no CPU fallback)")
    return t.float().contiguous()


# ------------------------------------------------------------------------------------------------ autograd functions
def _rope_nodes_(t: torch.Tensor, pos: torch.Tensor, head_dim: int, heads: int, m: int, base: float, inverse: bool) -> None:
    """Rotary embedding of per-node q / k rows ((N, head_dim, heads) layout, fp32) in place (layers.py:420-491)."""
    check(lib().gp_rope_nodes(C.c_void_p(ptr(t)), C.c_void_p(ptr(pos)), C.c_int32(pos.stride(0)), C.c_int64(t.shape[0]), C.c_int32(head_dim),
                              C.c_int32(heads), C.c_int32(pos.shape[1]), C.c_int32(m), C.c_float(base), C.c_int32(int(inverse)),
                              C.c_void_p(stream_ptr())), "gp_rope_nodes")
    ops._launched()


def _to_bf16(t: torch.Tensor) -> torch.Tensor:
    out = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
    ops.cast_bf16(t, out)
    return out


class _AttentionBranch(torch.autograd.Function):
    """out = [x +] proj(gate * attention(rope(q), rope(k), v)), q / k / v = Linear([RMSNorm](x)) (layers.py:637-697, 766-790);
    `rope` = (pos, m, base) or None, the gate (wg, bg) optional."""

    @staticmethod
    def forward(ctx, x, sn, wq, bq, wk, bk, wv, bv, wp, bp, wg, bg, add_resid: bool, g, heads: int, terms: int, rope):
        x = _f32(x)
        bf = terms == 1
        n = norm_fwd(x, sn, None, out_bf16=bf) if sn is not None else x
        H = wq.shape[0]
        fused = rope is None and _stackable((wq, wk, wv), (bq, bk, bv))
        if fused:                                # q | k | v = n [Wq; Wk; Wv]^T: ONE GEMM into one [N, 3H] buffer
            qkv = lin_fwd(n, _stack(wq, 3), _stack(bq, 3) if bq is not None else None, out_bf16=bf, terms=terms)
            q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
        elif rope is not None:
            pos, m, base = rope
            q = lin_fwd(n, wq, bq, terms=terms)
            k = lin_fwd(n, wk, bk, terms=terms)
            _rope_nodes_(q, pos, H // heads, heads, m, base, False)
            _rope_nodes_(k, pos, H // heads, heads, m, base, False)
            if bf:
                q, k = _to_bf16(q), _to_bf16(k)
        else:
            q = lin_fwd(n, wq, bq, out_bf16=bf, terms=terms)
            k = lin_fwd(n, wk, bk, out_bf16=bf, terms=terms)
        if not fused:
            v = lin_fwd(n, wv, bv, out_bf16=bf, terms=terms)
        y, y32, lse = attn_fwd(q, k, v, g, heads)
        gl = None
        yin = y
        if wg is not None:                       # gated attention (layers.py:684-689): y * sigmoid(gate_proj(x))
            gl = lin_fwd(n, wg, bg, terms=terms)
            ysrc = y32 if y32 is not None else y
            yin = torch.empty_like(ysrc)
            check(lib().gp_sigmoid_mul_fwd(C.c_void_p(ptr(gl)), C.c_void_p(ptr(ysrc)), C.c_int64(ysrc.numel()), C.c_void_p(ptr(yin)),
                                           C.c_void_p(stream_ptr())), "gp_sigmoid_mul_fwd")
            ops._launched()
        out = lin_fwd(yin, wp, bp, resid=x if add_resid else None, terms=terms)
        ctx.save_for_backward(x, sn, n if sn is not None else None, q, k, v, y, y32, lse, wq, wk, wv, wp, wg, gl, yin if wg is not None else None,
                              rope[0] if rope is not None else None)
        ctx.cfg = (add_resid, g, heads, terms, bq is not None, bk is not None, bv is not None, bp is not None, bg is not None,
                   (rope[1], rope[2]) if rope is not None else None, fused)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, sn, n, q, k, v, y, y32, lse, wq, wk, wv, wp, wg, gl, yin, pos = ctx.saved_tensors
        add_resid, g, heads, terms, hbq, hbk, hbv, hbp, hbg, rope, fused = ctx.cfg
        if n is None:
            n = x
        dout = _f32(dout)
        dy = lin_dgrad(dout, wp, terms=terms)
        proj_in = yin if wg is not None else y
        dwp, dbp = lin_wgrad(dout, proj_in, bias=True, terms=terms) if hbp else (lin_wgrad(dout, proj_in, terms=terms), None)
        dgl = dwg = dbg = None
        if wg is not None:
            ysrc = y32 if y32 is not None else y
            dgl, dy2 = torch.empty_like(gl), torch.empty_like(dy)
            check(lib().gp_sigmoid_mul_bwd(C.c_void_p(ptr(gl)), C.c_void_p(ptr(ysrc)), C.c_void_p(ptr(dy)), C.c_int64(dy.numel()),
                                           C.c_void_p(ptr(dgl)), C.c_void_p(ptr(dy2)), C.c_void_p(stream_ptr())), "gp_sigmoid_mul_bwd")
            ops._launched()
            dy = dy2
            dwg, dbg = lin_wgrad(dgl, n, bias=True, terms=terms) if hbg else (lin_wgrad(dgl, n, terms=terms), None)
        dq, dk, dv = attn_bwd(q, k, v, y, y32, lse, dy, g, heads, fused_out=fused)
        if fused:                                # one dgrad over K = 3H, one wgrad for [Wq; Wk; Wv] (+ the bias column)
            H = wq.shape[0]
            dqkv = dq._base if dq._base is not None else dq
            dn = lin_dgrad(dqkv, _stack(wq, 3), terms=terms)
            if dgl is not None:
                lin_dgrad(dgl, wg, out=dn, terms=terms)
            if hbq:
                dw, db = lin_wgrad(dqkv, n, bias=True, terms=terms)
                dbq, dbk, dbv = db[:H], db[H:2 * H], db[2 * H:]
            else:
                dw, dbq, dbk, dbv = lin_wgrad(dqkv, n, terms=terms), None, None, None
            dwq, dwk, dwv = dw[:H], dw[H:2 * H], dw[2 * H:]
            dsn = None
            if sn is not None:
                dx, dsn, _ = norm_bwd(x, sn, None, dn, add=dout if add_resid else None)
            else:
                assert not add_resid
                dx = dn
            return dx, dsn, dwq, dbq, dwk, dbk, dwv, dbv, dwp, dbp, dwg, dbg, None, None, None, None, None
        if rope is not None:                     # the rotation is orthogonal: its transpose is the inverse rotation
            H = wq.shape[0]
            _rope_nodes_(dq, pos, H // heads, heads, rope[0], rope[1], True)
            _rope_nodes_(dk, pos, H // heads, heads, rope[0], rope[1], True)
        dn = lin_dgrad(dq, wq, terms=terms)
        lin_dgrad(dk, wk, out=dn, terms=terms)
        lin_dgrad(dv, wv, out=dn, terms=terms)
        if dgl is not None:
            lin_dgrad(dgl, wg, out=dn, terms=terms)
        dwq, dbq = lin_wgrad(dq, n, bias=True, terms=terms) if hbq else (lin_wgrad(dq, n, terms=terms), None)
        dwk, dbk = lin_wgrad(dk, n, bias=True, terms=terms) if hbk else (lin_wgrad(dk, n, terms=terms), None)
        dwv, dbv = lin_wgrad(dv, n, bias=True, terms=terms) if hbv else (lin_wgrad(dv, n, terms=terms), None)
        dsn = None
        if sn is not None:
            dx, dsn, _ = norm_bwd(x, sn, None, dn, add=dout if add_resid else None)
        else:
            assert not add_resid
            dx = dn
        return dx, dsn, dwq, dbq, dwk, dbk, dwv, dbv, dwp, dbp, dwg, dbg, None, None, None, None, None


def attention_branch(x, norm_scale, att, g, add_resid: bool, terms: int = 1, pos=None):
    """`att`: a models.layers.Attention module (its q_proj / k_proj / v_proj / proj [/ gate_proj] parameters); `pos`: node
    positions when the module uses rotary embeddings."""
    rope = None
    if pos is not None and att.m > 0:
        rope = (pos[:, :att.pos_dimension].float().contiguous(), int(att.m), float(att.rope_base))
    wg, bg = (att.gate_proj.weight, att.gate_proj.bias) if att.gate_proj is not None else (None, None)
    return _AttentionBranch.apply(x, norm_scale, att.q_proj.weight, att.q_proj.bias, att.k_proj.weight, att.k_proj.bias,
                                  att.v_proj.weight, att.v_proj.bias, att.proj.weight, att.proj.bias, wg, bg, add_resid, g, att.num_heads,
                                  terms, rope)


class SparseAttention:
    """What `return_attention=True` hands back (the reference returns the dgl SparseMatrix of the softmax, layers.py:509-517,
    680-683): `val` [E, heads] in the order of the caller's edge_index, `row` / `col` = edge_index[0] / edge_index[1]."""

    def __init__(self, row: torch.Tensor, col: torch.Tensor, val: torch.Tensor, num_nodes: int):
        self.row, self.col, self.val, self.shape = row, col, val, (num_nodes, num_nodes)

    def coo(self):
        return self.row, self.col


@torch.no_grad()
def attention_weights(att, x, norm_scale, g, terms: int = 1, pos=None) -> SparseAttention:
    """softmax_j(q_i . k_j / sqrt(head_dim)) per stored entry (i, j) and head -- the attention matrix of `att` for input x
    (normalised first when `norm_scale` is given).  A diagnostic beside the fused path (which never materialises these
    values): projections on gp_gemm, RoPE by its kernel, the E x heads softmax with torch indexing in fp32."""
    x = _f32(x)
    n = norm_fwd(x, norm_scale, None, out_bf16=False) if norm_scale is not None else x
    heads, d = att.num_heads, att.head_dim
    q = lin_fwd(n, att.q_proj.weight, att.q_proj.bias, terms=terms)
    k = lin_fwd(n, att.k_proj.weight, att.k_proj.bias, terms=terms)
    if pos is not None and att.m > 0:
        p_ = pos[:, :att.pos_dimension].float().contiguous()
        _rope_nodes_(q, p_, d, heads, int(att.m), float(att.rope_base), False)
        _rope_nodes_(k, p_, d, heads, int(att.m), float(att.rope_base), False)
    row, col = g.src.long(), g.dst.long()                  # receiver-sorted positions; rows of the matrix = edge_index[0]
    E, N = row.numel(), n.shape[0]
    s = (q[row].view(E, d, heads) * k[col].view(E, d, heads)).sum(1) / math.sqrt(d)
    idx = row[:, None].expand(E, heads)
    mx = torch.full((N, heads), -float("inf"), dtype=s.dtype, device=s.device).scatter_reduce(0, idx, s, reduce="amax", include_self=True)
    pexp = torch.exp(s - mx[row])
    den = torch.zeros((N, heads), dtype=s.dtype, device=s.device).index_add_(0, row, pexp)
    a = pexp / den[row]
    perm = g.perm_dst64
    val = torch.empty_like(a)
    val[perm] = a                                          # back to the caller's edge order
    r0 = torch.empty_like(row); c0 = torch.empty_like(col)
    r0[perm] = row; c0[perm] = col
    return SparseAttention(r0, c0, val, N)


class _GatedBranch(torch.autograd.Function):
    """out = [x +] [W3 .] (GELU(W1 n + b1) * (W2 n + b2)), n = [norm(norm(x; s1); s2)] (layers.py:213-278, 791-819)."""

    @staticmethod
    def forward(ctx, x, s1, s2, w1, b1, w2, b2, w3, b3, add_resid: bool, terms: int, kind: int):
        x = _f32(x)
        bf = terms == 1
        n = norm_fwd(x, s1, s2, out_bf16=bf) if s1 is not None else x
        fused = _stackable((w1, w2), (b1, b2))
        if fused:                                # [a1 | a2] = n [W1; W2]^T: one GEMM, the gate reads the two column blocks
            G = w1.shape[0]
            a12 = lin_fwd(n, _stack(w1, 2), _stack(b1, 2) if b1 is not None else None, terms=terms)
            a1, a2 = a12[:, :G], a12[:, G:]
        else:
            a1 = lin_fwd(n, w1, b1, terms=terms)
            a2 = lin_fwd(n, w2, b2, terms=terms)
        gate = gelu_fwd(a1, a2, out_bf16=(bf and w3 is not None), kind=kind)
        out = lin_fwd(gate, w3, b3, resid=x if add_resid else None, terms=terms) if w3 is not None else gate
        ctx.save_for_backward(x, s1, s2, n if s1 is not None else None, a1, a2, gate if w3 is not None else None, w1, w2, w3)
        ctx.cfg = (add_resid, terms, b1 is not None, b2 is not None, b3 is not None, kind, fused)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, s1, s2, n, a1, a2, gate, w1, w2, w3 = ctx.saved_tensors
        add_resid, terms, hb1, hb2, hb3, kind, fused = ctx.cfg
        if n is None:
            n = x
        dout = _f32(dout)
        dw3 = db3 = None
        if w3 is not None:
            dg = lin_dgrad(dout, w3, terms=terms)
            dw3, db3 = lin_wgrad(dout, gate, bias=True, terms=terms) if hb3 else (lin_wgrad(dout, gate, terms=terms), None)
        else:
            dg = dout
        da1, da2 = gelu_bwd(a1, a2, dg, kind=kind, fused_out=fused)
        if fused:
            G = w1.shape[0]
            da12 = da1._base
            dn = lin_dgrad(da12, _stack(w1, 2), terms=terms)
            if hb1:
                dw, db = lin_wgrad(da12, n, bias=True, terms=terms)
                db1, db2 = db[:G], db[G:]
            else:
                dw, db1, db2 = lin_wgrad(da12, n, terms=terms), None, None
            dw1, dw2 = dw[:G], dw[G:]
        else:
            dn = lin_dgrad(da1, w1, terms=terms)
            lin_dgrad(da2, w2, out=dn, terms=terms)
            dw1, db1 = lin_wgrad(da1, n, bias=True, terms=terms) if hb1 else (lin_wgrad(da1, n, terms=terms), None)
            dw2, db2 = lin_wgrad(da2, n, bias=True, terms=terms) if hb2 else (lin_wgrad(da2, n, terms=terms), None)
        ds1 = ds2 = None
        if s1 is not None:
            dx, ds1, ds2 = norm_bwd(x, s1, s2, dn, add=dout if add_resid else None)
        else:
            assert not add_resid
            dx = dn
        return dx, ds1, ds2, dw1, db1, dw2, db2, dw3, db3, None, None, None


def gated_branch(x, s1, s2, gmlp, out_linear, add_resid: bool, terms: int = 1, act: str = "gelu"):
    """`gmlp`: a models.layers.GatedMLP; `out_linear`: the nn.Linear after it (None: return the gate itself, fp32); `act`:
    "gelu", or "silu" under the reference's global SiLU switch (layers.py:232-233)."""
    w3, b3 = (out_linear.weight, out_linear.bias) if out_linear is not None else (None, None)
    return _GatedBranch.apply(x, s1, s2, gmlp.linear1.weight, gmlp.linear1.bias, gmlp.linear2.weight, gmlp.linear2.bias, w3, b3,
                              add_resid, terms, 2 if act == "silu" else 3)


class _Mlp4(torch.autograd.Function):
    """build_mlp (layers.py:163-210): Linear, ReLU, Linear, ReLU, Linear, ReLU, Linear, [RMSNorm]; hidden activations are
    stored as bf16 (the next GEMM's operand; fp32 when terms = 3), the last Linear and the norm are fp32."""

    @staticmethod
    def forward(ctx, x, w0, b0, w1, b1, w2, b2, w3, b3, scale, terms: int):
        x = _f32(x)
        hb = terms == 1
        h0 = lin_fwd(x, w0, b0, relu=True, out_bf16=hb, terms=terms)
        h1 = lin_fwd(h0, w1, b1, relu=True, out_bf16=hb, terms=terms)
        h2 = lin_fwd(h1, w2, b2, relu=True, out_bf16=hb, terms=terms)
        y = lin_fwd(h2, w3, b3, terms=terms)
        out = norm_fwd(y, scale, None, out_bf16=False) if scale is not None else y
        ctx.save_for_backward(x, h0, h1, h2, y if scale is not None else None, scale, w0, w1, w2, w3)
        ctx.terms = terms
        return out

    @staticmethod
    def backward(ctx, dout):
        x, h0, h1, h2, y, scale, w0, w1, w2, w3 = ctx.saved_tensors
        terms = ctx.terms
        d = _f32(dout)
        dscale = None
        if scale is not None:
            d, dscale, _ = norm_bwd(y, scale, None, d)
        grads = []
        acts, ws = (x, h0, h1, h2), (w0, w1, w2, w3)
        for i in (3, 2, 1, 0):
            grads.append(lin_wgrad(d, acts[i], bias=True, terms=terms))
            if i > 0:
                d = lin_dgrad(d, ws[i], terms=terms)
                relu_mask_(d, acts[i])
        dx = lin_dgrad(d, w0, terms=terms) if ctx.needs_input_grad[0] else None
        (dw3, db3), (dw2, db2), (dw1, db1), (dw0, db0) = grads
        return dx, dw0, db0, dw1, db1, dw2, db2, dw3, db3, dscale, None


def mlp4(seq, x, terms: int = 1):
    scale = seq[7].scale if len(seq) > 7 else None
    return _Mlp4.apply(x, seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias, seq[4].weight, seq[4].bias, seq[6].weight,
                       seq[6].bias, scale, terms)


# small standalone functions (used by the unit tests and by RMSNorm-only callers); fp32 in, fp32 out under autograd
class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s1, s2, out_bf16: bool):
        x = _f32(x)
        ctx.save_for_backward(x, s1, s2)
        return norm_fwd(x, s1, s2, out_bf16)

    @staticmethod
    def backward(ctx, dy):
        x, s1, s2 = ctx.saved_tensors
        dx, ds1, ds2 = norm_bwd(x, s1, s2, _f32(dy))
        return dx, ds1, ds2, None


def rms_norm(x, s1, s2=None, out_bf16: bool = False):
    return _RMSNorm.apply(x, s1, s2, out_bf16)


class _GeluGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a1, a2, out_bf16: bool):
        a1, a2 = _f32(a1), _f32(a2)
        ctx.save_for_backward(a1, a2)
        return gelu_fwd(a1, a2, out_bf16)

    @staticmethod
    def backward(ctx, dg):
        a1, a2 = ctx.saved_tensors
        da1, da2 = gelu_bwd(a1, a2, _f32(dg))
        return da1, da2, None


def gelu_gate(a1, a2, out_bf16: bool = False):
    return _GeluGate.apply(a1, a2, out_bf16)
