"""ctypes binding of libgp_b200.so (the C ABI declared in include/gp_b200.h).

There is no CPU fallback: importing works without a GPU (so the C-ABI export test can run),
but every compute call goes to the CUDA library and raises if it is missing or reports an error.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GP_B200_LIB") or os.path.join(_HERE, "lib", "libgp_b200.so")   # override: A/B builds while tuning

_lib: Optional[C.CDLL] = None


class GpError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "graphphysics_b200 has no CPU or PyTorch fallback."
            )
        _lib = C.CDLL(LIB_PATH)
        _lib.gp_last_error.restype = C.c_char_p
        _lib.gp_version.restype = C.c_int
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise GpError(f"{what} failed ({rc}): {lib().gp_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "graphphysics_b200 kernels take CUDA tensors only"
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class MlpFwdArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("a_bf16", C.c_void_p),
        ("a_f32", C.c_void_p),
        ("ka", C.c_int32),
        ("lda", C.c_int32),
        ("init", C.c_void_p),
        ("ld_init", C.c_int32),
        ("init_off0", C.c_int32),
        ("init_off1", C.c_int32),
        ("idx0", C.c_void_p),
        ("idx1", C.c_void_p),
        ("two_inits", C.c_int32),
        ("n_layers", C.c_int32),
        ("w", C.c_void_p * 4),
        ("bias", C.c_void_p * 4),
        ("k", C.c_int32 * 4),
        ("n", C.c_int32 * 4),
        ("norm_scale", C.c_void_p),
        ("resid", C.c_void_p),
        ("y_bf16", C.c_void_p),
        ("y_f32", C.c_void_p),
        ("ld_out", C.c_int32),
        ("n_valid", C.c_int32),
        ("save_h2", C.c_void_p),
        ("save_h1", C.c_void_p),
        ("save_h3", C.c_void_p),
        ("seg_id", C.c_void_p),
        ("seg_out", C.c_void_p),
        ("seg_out_bf16", C.c_void_p),
        ("seg_bnd", C.c_void_p),
        ("prof", C.c_void_p),
    ]


class MlpBwdArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("a_bf16", C.c_void_p),
        ("a_f32", C.c_void_p),
        ("ka", C.c_int32),
        ("lda", C.c_int32),
        ("init", C.c_void_p),
        ("ld_init", C.c_int32),
        ("init_off0", C.c_int32),
        ("init_off1", C.c_int32),
        ("idx0", C.c_void_p),
        ("idx1", C.c_void_p),
        ("two_inits", C.c_int32),
        ("ha_saved", C.c_void_p),
        ("wa", C.c_void_p),
        ("ba", C.c_void_p),
        ("wb", C.c_void_p),
        ("bb", C.c_void_p),
        ("nb", C.c_int32),
        ("mode", C.c_int32),
        ("delta_b", C.c_void_p),
        ("ld_db", C.c_int32),
        ("norm_scale", C.c_void_p),
        ("gy_bf16", C.c_void_p),
        ("gy_f32", C.c_void_p),
        ("ld_gy", C.c_int32),
        ("gy_gather", C.c_void_p),
        ("gy_gather_bf16", C.c_void_p),
        ("gy_idx", C.c_void_p),
        ("need_din", C.c_int32),
        ("mask_by_ain", C.c_int32),
        ("out_resid", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("ld_out", C.c_int32),
        ("delta_a_out", C.c_void_p),
        ("seg_id", C.c_void_p),
        ("seg_out", C.c_void_p),
        ("seg_out_bf16", C.c_void_p),
        ("seg_bnd", C.c_void_p),
        ("partials", C.c_void_p),
        ("prof", C.c_void_p),
    ]


class LinearBwdArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("n_src", C.c_int32),
        ("src_f32", C.c_void_p * 3),
        ("src_bf16", C.c_void_p * 3),
        ("ld_src", C.c_int32 * 3),
        ("w", C.c_void_p),
        ("x", C.c_void_p),
        ("ldx", C.c_int32),
        ("dx_in", C.c_void_p),
        ("dx_out", C.c_void_p),
        ("partials", C.c_void_p),
    ]


class PackEntry(C.Structure):
    _fields_ = [
        ("src_off", C.c_int64),
        ("ld_src", C.c_int32),
        ("src_col0", C.c_int32),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("dst_off", C.c_int64),
        ("ld_dst", C.c_int32),
        ("dst_row0", C.c_int32),
        ("dst_col0", C.c_int32),
    ]


class ReduceSeg(C.Structure):
    _fields_ = [
        ("offset", C.c_int32),
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("ld_part", C.c_int32),
        ("dst", C.c_void_p),
        ("ld_dst", C.c_int32),
        ("accumulate", C.c_int32),
        ("partials", C.c_void_p),
        ("n_parts", C.c_int32),
        ("stride", C.c_int32),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("hidden", C.c_int32), ("num_heads", C.c_int32),
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("rowptr", C.c_void_p), ("col", C.c_void_p),
        ("y", C.c_void_p), ("lse", C.c_void_p),
        ("dy", C.c_void_p), ("pos", C.c_void_p), ("colptr", C.c_void_p), ("row", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("edge_a", C.c_void_p), ("edge_ds", C.c_void_p),
        ("io_bf16", C.c_int32),
        ("ld_qkv", C.c_int32), ("ld_dqkv", C.c_int32),
        ("y_f32", C.c_void_p),
    ]


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a", C.c_void_p), ("a_sm", C.c_int64), ("a_sk", C.c_int64),
        ("b", C.c_void_p), ("b_sn", C.c_int64), ("b_sk", C.c_int64),
        ("c", C.c_void_p), ("c_sm", C.c_int64), ("c_sn", C.c_int64),
        ("bias", C.c_void_p), ("resid", C.c_void_p),
        ("a_bf16", C.c_int32), ("b_bf16", C.c_int32), ("c_bf16", C.c_int32),
        ("relu", C.c_int32), ("accumulate", C.c_int32),
        ("terms", C.c_int32),
        ("split_k", C.c_int32),
        ("partials", C.c_void_p),
        ("flags", C.c_int32),
        ("b_ones", C.c_int32),
    ]


# header struct name -> ctypes class, beyond the core six (tests/test_boundary_cpu.py checks every one)
EXTRA_STRUCTS = {"gp_gemm_args": GemmArgs}
