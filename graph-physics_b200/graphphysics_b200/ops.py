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
