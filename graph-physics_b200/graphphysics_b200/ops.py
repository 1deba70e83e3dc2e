"""Thin Python wrappers over the C ABI: tensor-in / tensor-out, stream taken from PyTorch.

PyTorch is only the allocator and stream provider here; all arithmetic on the hot path runs in
libgp_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import MlpFwdArgs, check, lib, ptr, stream_ptr


COUNTERS = {"launches": 0}      # kernels of libgp_b200.so launched through these wrappers


class _Profile:
    """Opt-in CUDA-event timing of selected kernel tags on the launching stream (bench.py)."""

    def __init__(self):
        self.tags, self.records = (), {}

    def reset(self, tags=()):
        self.tags, self.records = tuple(tags), {t: [] for t in tags}

    def begin(self, tag):
        if tag is None or tag not in self.tags:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def end(self, tag, e0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records[tag].append((e0, e1))

    def summary(self):
        torch.cuda.synchronize()
        return {t: {"total_ms": sum(a.elapsed_time(b) for a, b in r), "calls": len(r)}
                for t, r in self.records.items() if r}


PROFILE = _Profile()


def _launched(n: int = 1) -> None:
    COUNTERS["launches"] += n


def pad16(v: int) -> int:
    return (v + 15) // 16 * 16


def pack_weight(w: torch.Tensor, n_pad: Optional[int] = None, k_pad: Optional[int] = None) -> torch.Tensor:
    """fp32 nn.Linear weight (out, in) -> packed bf16 [pad16(out)][pad16(in)], zero padded."""
    n, k = w.shape
    n_pad = n_pad or pad16(n)
    k_pad = k_pad or pad16(k)
    out = torch.zeros((n_pad, k_pad), dtype=torch.bfloat16, device=w.device)
    out[:n, :k] = w.detach().to(torch.bfloat16)
    return out


def pack_bias(b: Optional[torch.Tensor], n_pad: int, device) -> torch.Tensor:
    out = torch.zeros((n_pad,), dtype=torch.float32, device=device)
    if b is not None:
        out[: b.numel()] = b.detach().float()
    return out


def seg_sub_rows(hidden: int, backward: bool = False) -> int:
    """Rows per sub-tile of a kernel's segment walk (gp_seg_sub_rows)."""
    return int(lib().gp_seg_sub_rows(C.c_int32(hidden), C.c_int32(1 if backward else 0)))


def seg_bnd_size(rows: int, hidden: int, backward: bool = False) -> int:
    sub = seg_sub_rows(hidden, backward)
    return ((rows + sub - 1) // sub) * 2 * hidden


def mlp_fwd(
    rows: int,
    hidden: int,
    weights: Sequence[torch.Tensor],
    biases: Sequence[Optional[torch.Tensor]],
    *,
    a: torch.Tensor,
    ka: int,
    init: Optional[torch.Tensor] = None,
    init_off0: int = 0,
    init_off1: int = 0,
    idx0: Optional[torch.Tensor] = None,
    idx1: Optional[torch.Tensor] = None,
    two_inits: bool = False,
    norm_scale: Optional[torch.Tensor] = None,
    resid: Optional[torch.Tensor] = None,
    out: torch.Tensor,
    n_valid: int,
    save_h2: Optional[torch.Tensor] = None,
    save_h1: Optional[torch.Tensor] = None,
    save_h3: Optional[torch.Tensor] = None,
    seg_id: Optional[torch.Tensor] = None,
    seg_out: Optional[torch.Tensor] = None,
    seg_bnd: Optional[torch.Tensor] = None,
    tag: Optional[str] = None,
    prof: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """gp_mlp_fwd (include/gp_b200.h).  `weights` are packed bf16 [n][k]; `a` is [rows, >=ka]
    bf16 or fp32 with unit column stride; `out` is [rows, ld] bf16 or fp32."""
    args = MlpFwdArgs()
    args.rows = rows
    if a.dtype == torch.bfloat16:
        args.a_bf16 = ptr(a)
    else:
        assert a.dtype == torch.float32
        args.a_f32 = ptr(a)
    assert a.stride(-1) == 1
    args.ka, args.lda = ka, a.stride(0)
    if init is not None:
        assert init.dtype == torch.bfloat16 and init.stride(-1) == 1
        args.init, args.ld_init = ptr(init), init.stride(0)
        args.init_off0, args.init_off1 = init_off0, init_off1
        args.idx0, args.idx1 = ptr(idx0), ptr(idx1)
        args.two_inits = 1 if two_inits else 0
    args.n_layers = len(weights)
    for l, (w, b) in enumerate(zip(weights, biases)):
        assert w.dtype == torch.bfloat16 and w.is_contiguous()
        args.w[l] = ptr(w)
        args.bias[l] = ptr(b)
        args.n[l], args.k[l] = w.shape
    args.norm_scale = ptr(norm_scale)
    args.resid = ptr(resid)
    if out.dtype == torch.bfloat16:
        args.y_bf16 = ptr(out)
    else:
        assert out.dtype == torch.float32
        args.y_f32 = ptr(out)
    args.ld_out = out.stride(0)
    args.n_valid = n_valid
    args.save_h2, args.save_h1, args.save_h3 = ptr(save_h2), ptr(save_h1), ptr(save_h3)
    for sv in (save_h1, save_h2, save_h3):
        assert sv is None or (sv.dtype == torch.bfloat16 and sv.is_contiguous() and sv.shape == (rows, hidden))
    args.seg_id, args.seg_bnd = ptr(seg_id), ptr(seg_bnd)
    if seg_out is not None and seg_out.dtype == torch.bfloat16:
        args.seg_out_bf16 = ptr(seg_out)
    else:
        args.seg_out = ptr(seg_out)
    args.prof = ptr(prof)
    ev = PROFILE.begin(tag)
    check(lib().gp_mlp_fwd(C.byref(args), C.c_int(hidden), C.c_void_p(stream_ptr())), "gp_mlp_fwd")
    PROFILE.end(tag, ev)
    _launched()
    return out


def set_launch_overlap(enabled: bool) -> bool:
    """gp_set_launch_overlap: programmatic dependent launch for the persistent MLP kernels (their weight-staging
    prologue overlaps the tail of the previous kernel).  Safe when the packed weights / biases are not written
    by the launch immediately before an MLP kernel, which holds for the engine.  Returns the previous setting."""
    return bool(lib().gp_set_launch_overlap(1 if enabled else 0))


def seg_fixup(rowptr: torch.Tensor, hidden: int, seg_bnd: torch.Tensor, seg_out: torch.Tensor,
              backward: bool = False) -> None:
    assert rowptr.dtype == torch.int32
    fn = lib().gp_seg_fixup_bf16 if seg_out.dtype == torch.bfloat16 else lib().gp_seg_fixup
    check(
        fn(C.c_void_p(ptr(rowptr)), C.c_int32(rowptr.numel() - 1), C.c_int32(hidden),
                           C.c_int32(seg_sub_rows(hidden, backward)), C.c_void_p(ptr(seg_bnd)), C.c_void_p(ptr(seg_out)),
                           C.c_void_p(stream_ptr())),
        "gp_seg_fixup",
    )
    _launched()


from ._lib import MlpBwdArgs  # noqa: E402

_SM_COUNT = None


def sm_count() -> int:
    global _SM_COUNT
    if _SM_COUNT is None:
        out = (C.c_int * 3)()
        check(lib().gp_device_info(out), "gp_device_info")
        _SM_COUNT = int(out[0])
    return _SM_COUNT


def bwd_layout(hidden: int, ka: int, nb: int):
    """(off_dWb, off_dWa, off_dbb, off_dba, off_dscale, stride) in floats."""
    out = (C.c_int32 * 6)()
    check(lib().gp_mlp_bwd_layout(hidden, ka, nb, out), "gp_mlp_bwd_layout")
    return tuple(int(v) for v in out)


def mlp_bwd_stage(
    rows: int,
    hidden: int,
    *,
    a: torch.Tensor,
    ka: int,
    wa: torch.Tensor,
    ba: Optional[torch.Tensor],
    wb: torch.Tensor,
    bb: Optional[torch.Tensor],
    partials: torch.Tensor,
    init: Optional[torch.Tensor] = None,
    init_off0: int = 0,
    init_off1: int = 0,
    idx0: Optional[torch.Tensor] = None,
    idx1: Optional[torch.Tensor] = None,
    two_inits: bool = False,
    ha_saved: Optional[torch.Tensor] = None,
    delta_b: Optional[torch.Tensor] = None,
    norm_scale: Optional[torch.Tensor] = None,
    gy: Optional[torch.Tensor] = None,
    gy_gather: Optional[torch.Tensor] = None,
    gy_idx: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
    mask_by_ain: bool = False,
    out_resid: Optional[torch.Tensor] = None,
    delta_a_out: Optional[torch.Tensor] = None,
    seg_id: Optional[torch.Tensor] = None,
    seg_out: Optional[torch.Tensor] = None,
    seg_bnd: Optional[torch.Tensor] = None,
    tag: Optional[str] = None,
    prof: Optional[torch.Tensor] = None,
) -> int:
    """gp_mlp_bwd_stage (include/gp_b200.h).  NORM mode when `gy` is given, GIVEN mode when
    `delta_b` is given.  Returns the number of partial blocks written to `partials`."""
    args = MlpBwdArgs()
    args.rows = rows
    if a.dtype == torch.bfloat16:
        args.a_bf16 = ptr(a)
    else:
        assert a.dtype == torch.float32
        args.a_f32 = ptr(a)
    args.ka, args.lda = ka, a.stride(0)
    if init is not None:
        args.init, args.ld_init = ptr(init), init.stride(0)
        args.init_off0, args.init_off1 = init_off0, init_off1
        args.idx0, args.idx1 = ptr(idx0), ptr(idx1)
        args.two_inits = 1 if two_inits else 0
    if ha_saved is not None:
        assert ha_saved.dtype == torch.bfloat16 and ha_saved.is_contiguous() and ha_saved.shape == (rows, hidden)
        args.ha_saved = ptr(ha_saved)
    args.wa, args.ba, args.wb, args.bb = ptr(wa), ptr(ba), ptr(wb), ptr(bb)
    args.nb = wb.shape[0]
    assert wa.shape == (hidden, ka) and wb.shape[1] == hidden
    if delta_b is None:
        args.mode = 1
        args.norm_scale = ptr(norm_scale)
        if gy is not None:
            if gy.dtype == torch.bfloat16:
                args.gy_bf16 = ptr(gy)
            else:
                assert gy.dtype == torch.float32
                args.gy_f32 = ptr(gy)
            args.ld_gy = gy.stride(0)
        if gy_gather is not None and gy_gather.dtype == torch.bfloat16:
            assert gy_gather.is_contiguous() and gy_gather.shape[1] == hidden
            args.gy_gather_bf16 = ptr(gy_gather)
        else:
            args.gy_gather = ptr(gy_gather)
        args.gy_idx = ptr(gy_idx)
    else:
        args.mode = 0
        assert delta_b is not None and delta_b.dtype == torch.bfloat16
        args.delta_b, args.ld_db = ptr(delta_b), delta_b.stride(0)
    if out is not None:
        args.need_din = 1
        if out.dtype == torch.bfloat16:
            args.out_bf16 = ptr(out)
        else:
            assert out.dtype == torch.float32
            args.out_f32 = ptr(out)
        args.ld_out = out.stride(0)
        args.mask_by_ain = 1 if mask_by_ain else 0
        args.out_resid = ptr(out_resid)
    args.delta_a_out = ptr(delta_a_out)
    args.seg_id, args.seg_bnd = ptr(seg_id), ptr(seg_bnd)
    if seg_out is not None and seg_out.dtype == torch.bfloat16:
        args.seg_out_bf16 = ptr(seg_out)
    else:
        args.seg_out = ptr(seg_out)
    stride = bwd_layout(hidden, ka, args.nb)[5]
    assert partials.dtype == torch.float32 and partials.numel() >= stride * min(sm_count(), (rows + 127) // 128)
    args.partials = ptr(partials)
    args.prof = ptr(prof)
    grid = C.c_int32(0)
    ev = PROFILE.begin(tag)
    check(lib().gp_mlp_bwd_stage(C.byref(args), C.c_int(hidden), C.byref(grid), C.c_void_p(stream_ptr())),
          "gp_mlp_bwd_stage")
    PROFILE.end(tag, ev)
    _launched()
    return int(grid.value)


def reduce_partials(partials: torch.Tensor, n_parts: int, stride: int, offset: int, rows: int, cols: int, ld_part: int,
                    dst: torch.Tensor, ld_dst: int, accumulate: bool) -> None:
    check(
        lib().gp_reduce_partials(C.c_void_p(ptr(partials)), n_parts, stride, offset, rows, cols, ld_part,
                                 C.c_void_p(ptr(dst)), ld_dst, 1 if accumulate else 0, C.c_void_p(stream_ptr())),
        "gp_reduce_partials",
    )
    _launched()


from ._lib import LinearBwdArgs, PackEntry, ReduceSeg  # noqa: E402


def linear_bwd(rows: int, hidden: int, srcs: Sequence[torch.Tensor], w: torch.Tensor, x: torch.Tensor,
               dx_in: Optional[torch.Tensor], dx_out: torch.Tensor, partials: torch.Tensor) -> int:
    """gp_linear_bwd: dx_out = dx_in + sum_s srcs[s] . W_s ;  per-CTA dW partials."""
    args = LinearBwdArgs()
    args.rows, args.n_src = rows, len(srcs)
    for s, t in enumerate(srcs):
        if t.dtype == torch.float32:
            args.src_f32[s] = ptr(t)
        else:
            assert t.dtype == torch.bfloat16
            args.src_bf16[s] = ptr(t)
        args.ld_src[s] = t.stride(0)
    args.w, args.x, args.ldx = ptr(w), ptr(x), x.stride(0)
    args.dx_in, args.dx_out, args.partials = ptr(dx_in), ptr(dx_out), ptr(partials)
    grid = C.c_int32(0)
    check(lib().gp_linear_bwd(C.byref(args), C.c_int(hidden), C.byref(grid), C.c_void_p(stream_ptr())), "gp_linear_bwd")
    _launched()
    return int(grid.value)


def segsum_gather(src: torch.Tensor, perm: Optional[torch.Tensor], rowptr: torch.Tensor, hidden: int,
                  out: torch.Tensor) -> None:
    fn = lib().gp_segsum_gather_bf16 if out.dtype == torch.bfloat16 else lib().gp_segsum_gather
    check(
        fn(C.c_void_p(ptr(src)), C.c_int32(src.stride(0)), C.c_void_p(ptr(perm)),
                               C.c_void_p(ptr(rowptr)), C.c_int32(rowptr.numel() - 1), C.c_int32(hidden),
                               C.c_void_p(ptr(out)), C.c_void_p(stream_ptr())),
        "gp_segsum_gather",
    )
    _launched()


def reduce_multi(partials: Optional[torch.Tensor], n_parts: int, stride: int, segs) -> None:
    """segs: list of (offset, rows, cols, ld_part, dst_ptr(int), ld_dst, accumulate) -- reduced from the
    call-level `partials` (n_parts blocks of `stride` floats) -- optionally extended by
    (partials_ptr(int), n_parts, stride) for a segment that lives in another partial buffer.
    At most 32 segments per launch (longer lists are split; one processor layer's backward queues 21)."""
    for lo in range(0, len(segs), 32):
        chunk = segs[lo:lo + 32]
        arr = (ReduceSeg * len(chunk))()
        for i, sg in enumerate(chunk):
            off, rows, cols, ldp, dst, ldd, acc = sg[:7]
            arr[i].offset, arr[i].rows, arr[i].cols, arr[i].ld_part = off, rows, cols, ldp
            arr[i].dst, arr[i].ld_dst, arr[i].accumulate = dst, ldd, 1 if acc else 0
            if len(sg) > 7:
                arr[i].partials, arr[i].n_parts, arr[i].stride = sg[7], sg[8], sg[9]
            else:
                arr[i].partials, arr[i].n_parts, arr[i].stride = None, 0, 0
        check(lib().gp_reduce_partials_multi(C.c_void_p(ptr(partials) if partials is not None else None), n_parts, stride,
                                             arr, len(chunk), C.c_void_p(stream_ptr())), "gp_reduce_partials_multi")
        _launched()


_MSE_WS: dict = {}


def masked_mse(out: torch.Tensor, target: torch.Tensor, mask_u8: torch.Tensor, loss: torch.Tensor,
               grad: Optional[torch.Tensor], grad_scale: float = 1.0) -> None:
    n, d = out.shape
    assert out.is_contiguous() and target.is_contiguous() and mask_u8.dtype == torch.uint8
    ws = _MSE_WS.get(out.device)
    if ws is None:
        ws = _MSE_WS[out.device] = torch.empty(260, dtype=torch.float32, device=out.device)
    check(lib().gp_masked_mse(C.c_void_p(ptr(out)), C.c_void_p(ptr(target)), C.c_void_p(ptr(mask_u8)), n, d,
                              C.c_void_p(ptr(loss)), C.c_void_p(ptr(grad)), C.c_float(grad_scale),
                              C.c_void_p(ptr(ws)), C.c_void_p(stream_ptr())), "gp_masked_mse")
    _launched(3 if grad is not None else 2)


def sqnorm(g: torch.Tensor, workspace: torch.Tensor, out: torch.Tensor) -> None:
    check(lib().gp_sqnorm(C.c_void_p(ptr(g)), C.c_int64(g.numel()), C.c_void_p(ptr(workspace)), C.c_void_p(ptr(out)),
                          C.c_void_p(stream_ptr())), "gp_sqnorm")
    _launched(2)


def adamw(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, max_norm, sqnorm_t) -> None:
    check(lib().gp_adamw(C.c_void_p(ptr(params)), C.c_void_p(ptr(grads)), C.c_void_p(ptr(exp_avg)),
                         C.c_void_p(ptr(exp_avg_sq)), C.c_int64(params.numel()), C.c_float(lr), C.c_float(beta1),
                         C.c_float(beta2), C.c_float(eps), C.c_float(weight_decay), C.c_int32(step), C.c_float(max_norm),
                         C.c_void_p(ptr(sqnorm_t)), C.c_void_p(stream_ptr())), "gp_adamw")
    _launched()


def adamw_sched(params, grads, exp_avg, exp_avg_sq, state, base_lr, warmup, max_iters, min_lr_factor, beta1, beta2, eps,
                weight_decay, max_norm, sqnorm_t) -> None:
    check(lib().gp_adamw_sched(C.c_void_p(ptr(params)), C.c_void_p(ptr(grads)), C.c_void_p(ptr(exp_avg)),
                               C.c_void_p(ptr(exp_avg_sq)), C.c_int64(params.numel()), C.c_void_p(ptr(state)),
                               C.c_float(base_lr), C.c_int32(warmup), C.c_int32(max_iters), C.c_float(min_lr_factor),
                               C.c_float(beta1), C.c_float(beta2), C.c_float(eps), C.c_float(weight_decay),
                               C.c_float(max_norm), C.c_void_p(ptr(sqnorm_t)), C.c_void_p(stream_ptr())), "gp_adamw_sched")
    _launched(2)


def pack_weights(params: torch.Tensor, packed: torch.Tensor, table_dev: torch.Tensor, n_entries: int) -> None:
    check(lib().gp_pack_weights(C.c_void_p(ptr(params)), C.c_void_p(ptr(packed)), C.c_void_p(ptr(table_dev)),
                                C.c_int32(n_entries), C.c_void_p(stream_ptr())), "gp_pack_weights")
    _launched()


def cast_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    check(lib().gp_cast_bf16(C.c_void_p(ptr(src)), C.c_void_p(ptr(dst)), C.c_int64(src.numel()),
                             C.c_void_p(stream_ptr())), "gp_cast_bf16")
    _launched()


from ._lib import AttentionArgs  # noqa: E402


class CSRAttention(torch.autograd.Function):
    """y = masked multi-head attention(q, k, v) over the adjacency of `g` (rows = edge_index[0],
    columns = edge_index[1]); gp_csr_attention_fwd / _bwd.  q, k, v: [N, H] contiguous, all fp32 or all bf16
    (bf16: half the gathered bytes; y is then bf16 too, and an unrounded copy is kept for the backward)."""

    @staticmethod
    def forward(ctx, q, k, v, g, num_heads: int):
        from . import dense
        bf = q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16
        if bf:
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        else:
            q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
        y, y32, lse = dense.attn_fwd(q, k, v, g, num_heads)
        ctx.save_for_backward(q, k, v, y, y32, lse)
        ctx.g, ctx.num_heads = g, num_heads
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import dense
        q, k, v, y, y32, lse = ctx.saved_tensors
        dq, dk, dv = dense.attn_bwd(q, k, v, y, y32, lse, dy.contiguous().float(), ctx.g, ctx.num_heads)
        return dq, dk, dv, None, None


# ---------------------------------------------------------------------------------------------- graph layout / halo rows
def csr_from_coo(edge_index: torch.Tensor, num_nodes: int, out: dict, workspace: torch.Tensor,
                 prev: Optional[torch.Tensor] = None, state: Optional[torch.Tensor] = None) -> None:
    """gp_csr_from_coo: receiver-sorted layout of a CUDA edge_index [2, E] int64 into the int32 tensors of `out`
    (perm_dst, src, dst, rowptr_dst, perm_src, rowptr_src, att_col).  prev/state: skip when the topology is unchanged."""
    assert edge_index.is_cuda and edge_index.dtype == torch.int64 and edge_index.is_contiguous()
    lib().gp_csr_from_coo.restype = C.c_int
    check(lib().gp_csr_from_coo(C.c_void_p(ptr(edge_index)), C.c_int64(edge_index.shape[1]), C.c_int32(num_nodes),
                                C.c_void_p(ptr(out["perm_dst"])), C.c_void_p(ptr(out["src"])), C.c_void_p(ptr(out["dst"])),
                                C.c_void_p(ptr(out["rowptr_dst"])), C.c_void_p(ptr(out["perm_src"])),
                                C.c_void_p(ptr(out["rowptr_src"])), C.c_void_p(ptr(out["att_col"])), C.c_void_p(ptr(workspace)),
                                C.c_void_p(ptr(prev)), C.c_void_p(ptr(state)), C.c_void_p(stream_ptr())), "gp_csr_from_coo")
    _launched(21)


def csr_workspace_bytes(num_edges: int, num_nodes: int) -> int:
    lib().gp_csr_workspace_bytes.restype = C.c_int64
    return int(lib().gp_csr_workspace_bytes(C.c_int64(num_edges), C.c_int32(num_nodes)))


def halo_pack(x: torch.Tensor, idx: torch.Tensor, out: torch.Tensor) -> None:
    """out[r] = x[idx[r]] (gp_halo_pack)."""
    assert x.is_cuda and x.stride(1) == 1 and idx.dtype == torch.int32 and out.is_contiguous() and out.dtype == x.dtype
    check(lib().gp_halo_pack(C.c_void_p(ptr(x)), C.c_int32(x.stride(0)), C.c_int32(x.element_size()), C.c_void_p(ptr(idx)),
                             C.c_int32(idx.numel()), C.c_int32(x.shape[1]), C.c_void_p(ptr(out)), C.c_void_p(stream_ptr())), "gp_halo_pack")
    _launched()


def halo_unpack(x: torch.Tensor, idx: torch.Tensor, rows: torch.Tensor) -> None:
    """x[idx[r]] = rows[r] (gp_halo_unpack)."""
    assert x.is_cuda and x.stride(1) == 1 and idx.dtype == torch.int32 and rows.is_contiguous() and rows.dtype == x.dtype
    check(lib().gp_halo_unpack(C.c_void_p(ptr(x)), C.c_int32(x.stride(0)), C.c_int32(x.element_size()), C.c_void_p(ptr(idx)),
                               C.c_int32(idx.numel()), C.c_int32(x.shape[1]), C.c_void_p(ptr(rows)), C.c_void_p(stream_ptr())),
          "gp_halo_unpack")
    _launched()


def halo_unpack_add(x: torch.Tensor, dst_rows: torch.Tensor, rowptr: torch.Tensor, order: torch.Tensor, rows: torch.Tensor) -> None:
    """x[dst_rows[d]] += sum of rows[order[j]], j in [rowptr[d], rowptr[d+1])  (gp_halo_unpack_add, fp32, fixed order)."""
    assert x.dtype == torch.float32 and rows.dtype == torch.float32 and x.stride(1) == 1 and rows.is_contiguous()
    check(lib().gp_halo_unpack_add(C.c_void_p(ptr(x)), C.c_int32(x.stride(0)), C.c_void_p(ptr(dst_rows)), C.c_void_p(ptr(rowptr)),
                                   C.c_void_p(ptr(order)), C.c_int32(dst_rows.numel()), C.c_int32(x.shape[1]), C.c_void_p(ptr(rows)),
                                   C.c_void_p(stream_ptr())), "gp_halo_unpack_add")
    _launched()


# ---------------------------------------------------------------------------------------------- normaliser / node features
def normalizer_stats(x: torch.Tensor) -> torch.Tensor:
    """Per-block column sums of x and x^2 (gp_normalizer_stats): [blocks, 2 * size] fp32."""
    lib().gp_normalizer_blocks.restype = C.c_int32
    part = torch.empty((lib().gp_normalizer_blocks(), 2 * x.shape[1]), dtype=torch.float32, device=x.device)
    check(lib().gp_normalizer_stats(C.c_void_p(ptr(x)), C.c_int64(x.shape[0]), C.c_int32(x.shape[1]), C.c_int64(x.stride(0)),
                                    C.c_void_p(ptr(part)), C.c_void_p(stream_ptr())), "gp_normalizer_stats")
    _launched()
    return part


def normalizer_update(part: torch.Tensor, rows: int, norm=None, want_stats: bool = False) -> Optional[torch.Tensor]:
    """Sum the block partials; with `norm` also Normalizer._accumulate (gated on the device); with want_stats return
    [sum | sum of squares | rows]."""
    size = part.shape[1] // 2
    stats = torch.empty(2 * size + 1, dtype=torch.float32, device=part.device) if want_stats else None
    acc = (norm._acc_sum, norm._acc_sum_squared, norm._acc_count, norm._num_accumulations) if norm is not None else (None,) * 4
    check(lib().gp_normalizer_update(C.c_void_p(ptr(part)), C.c_int32(size), C.c_int64(rows), C.c_void_p(ptr(stats)),
                                     *(C.c_void_p(ptr(t)) for t in acc), C.c_float(float(norm._max_accumulations) if norm is not None else 0.0),
                                     C.c_void_p(stream_ptr())), "gp_normalizer_update")
    _launched()
    return stats


def normalizer_accumulate(stats: torch.Tensor, norm) -> None:
    """Normalizer._accumulate from a [sum | sum of squares | rows] vector (gp_normalizer_accumulate)."""
    size = (stats.numel() - 1) // 2
    check(lib().gp_normalizer_accumulate(C.c_void_p(ptr(stats)), C.c_int32(size), C.c_void_p(ptr(norm._acc_sum)),
                                         C.c_void_p(ptr(norm._acc_sum_squared)), C.c_void_p(ptr(norm._acc_count)),
                                         C.c_void_p(ptr(norm._num_accumulations)), C.c_float(float(norm._max_accumulations)),
                                         C.c_void_p(stream_ptr())), "gp_normalizer_accumulate")
    _launched()


def normalizer_apply(norm, x: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    """(x - mean) / max(std, eps) or its inverse from the accumulators of `norm` (gp_normalizer_apply)."""
    out = torch.empty((x.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    check(lib().gp_normalizer_apply(C.c_void_p(ptr(x)), C.c_int64(x.shape[0]), C.c_int32(x.shape[1]), C.c_int64(x.stride(0)),
                                    C.c_void_p(ptr(norm._acc_sum)), C.c_void_p(ptr(norm._acc_sum_squared)), C.c_void_p(ptr(norm._acc_count)),
                                    C.c_float(norm._eps_host), C.c_int32(int(inverse)), C.c_void_p(ptr(out)), C.c_int64(out.stride(0)),
                                    C.c_void_p(stream_ptr())), "gp_normalizer_apply")
    _launched()
    return out


def node_features(x: torch.Tensor, f0: int, f1: int, type_col: int, num_types: int) -> torch.Tensor:
    """[x[:, f0:f1] | one_hot(x[:, type_col], num_types)] in one launch (gp_node_features)."""
    out = torch.empty((x.shape[0], (f1 - f0) + num_types), dtype=torch.float32, device=x.device)
    check(lib().gp_node_features(C.c_void_p(ptr(x)), C.c_int64(x.shape[0]), C.c_int64(x.stride(0)), C.c_int32(f0), C.c_int32(f1),
                                 C.c_int32(type_col), C.c_int32(num_types), C.c_void_p(ptr(out)), C.c_void_p(stream_ptr())), "gp_node_features")
    _launched()
    return out
