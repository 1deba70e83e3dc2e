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


def seg_bnd_size(rows: int, hidden: int) -> int:
    sub = hidden // 2
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
    seg_id: Optional[torch.Tensor] = None,
    seg_out: Optional[torch.Tensor] = None,
    seg_bnd: Optional[torch.Tensor] = None,
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
    args.save_h2 = ptr(save_h2)
    args.seg_id, args.seg_out, args.seg_bnd = ptr(seg_id), ptr(seg_out), ptr(seg_bnd)
    check(lib().gp_mlp_fwd(C.byref(args), C.c_int(hidden), C.c_void_p(stream_ptr())), "gp_mlp_fwd")
    return out


def seg_fixup(rowptr: torch.Tensor, hidden: int, seg_bnd: torch.Tensor, seg_out: torch.Tensor) -> None:
    assert rowptr.dtype == torch.int32
    check(
        lib().gp_seg_fixup(C.c_void_p(ptr(rowptr)), C.c_int32(rowptr.numel() - 1), C.c_int32(hidden),
                           C.c_void_p(ptr(seg_bnd)), C.c_void_p(ptr(seg_out)), C.c_void_p(stream_ptr())),
        "gp_seg_fixup",
    )


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
    args.wa, args.ba, args.wb, args.bb = ptr(wa), ptr(ba), ptr(wb), ptr(bb)
    args.nb = wb.shape[0]
    assert wa.shape == (hidden, ka) and wb.shape[1] == hidden
    if gy is not None:
        args.mode = 1
        args.norm_scale = ptr(norm_scale)
        if gy.dtype == torch.bfloat16:
            args.gy_bf16 = ptr(gy)
        else:
            assert gy.dtype == torch.float32
            args.gy_f32 = ptr(gy)
        args.ld_gy = gy.stride(0)
        args.gy_gather, args.gy_idx = ptr(gy_gather), ptr(gy_idx)
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
    args.seg_id, args.seg_out, args.seg_bnd = ptr(seg_id), ptr(seg_out), ptr(seg_bnd)
    stride = bwd_layout(hidden, ka, args.nb)[5]
    assert partials.dtype == torch.float32 and partials.numel() >= stride * min(sm_count(), (rows + 127) // 128)
    args.partials = ptr(partials)
    grid = C.c_int32(0)
    check(lib().gp_mlp_bwd_stage(C.byref(args), C.c_int(hidden), C.byref(grid), C.c_void_p(stream_ptr())),
          "gp_mlp_bwd_stage")
    return int(grid.value)


def reduce_partials(partials: torch.Tensor, n_parts: int, stride: int, offset: int, rows: int, cols: int, ld_part: int,
                    dst: torch.Tensor, ld_dst: int, accumulate: bool) -> None:
    check(
        lib().gp_reduce_partials(C.c_void_p(ptr(partials)), n_parts, stride, offset, rows, cols, ld_part,
                                 C.c_void_p(ptr(dst)), ld_dst, 1 if accumulate else 0, C.c_void_p(stream_ptr())),
        "gp_reduce_partials",
    )
