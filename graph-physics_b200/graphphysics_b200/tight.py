"""precision="tight": the encode-process-decode model evaluated with split-precision tensor-core GEMMs.

Every dense contraction -- forward, dgrad and wgrad of every nn.Linear of build_mlp
(graphphysics/models/layers.py:163-210) -- runs in gp_gemm with terms = 3 (csrc/gemm.cu): fp32 operands split into three bf16
terms (hi + mid + lo, 24 mantissa bits), six tcgen05 MMAs per product, fp32 accumulate, fp32 tensors in HBM.  The algorithm is the one the
fused bf16 kernels execute: the node-dependent column blocks of both first layers are applied once
per node (P = x.[W1d; W1s; W1x]^T) and gathered as pre-activations.  Bias, ReLU, RMSNorm, gathers, the receiver sum
and the residuals are fp32 elementwise PyTorch ops under autograd, so this mode is a VERIFICATION mode, not the
timed path: it shows the model within 1e-3 (in fact ~1e-5) of the fp32 reference end to end -- one-step
prediction, loss and gradients -- which no bf16-operand evaluation of a 15-layer residual stack can
(SURVEY Appendix B).  Select it with EncodeProcessDecode(..., precision="tight"), `"precision": "tight"` in the
model section of the training JSON, or GP_B200_PRECISION=tight.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .dense import gemm as _gemm


def gemm3(M: int, N: int, K: int, a: torch.Tensor, a_sm: int, a_sk: int, b: torch.Tensor, b_sn: int, b_sk: int,
          c: torch.Tensor, c_sm: int, c_sn: int, bias: Optional[torch.Tensor] = None, relu: bool = False,
          accumulate: bool = False, split_k: int = 1) -> None:
    """C(m,n) = [C +] bias[n] + sum_k A(m,k) B(n,k) with strided fp32 operands, three-term split (gp_gemm, terms = 3)."""
    assert a.dtype == b.dtype == c.dtype == torch.float32
    _gemm(M, N, K, a, a_sm, a_sk, b, b_sn, b_sk, c, c_sm, c_sn, bias=bias, relu=relu, accumulate=accumulate, split_k=split_k, terms=3)


class _Linear3(torch.autograd.Function):
    """y = x W^T + b with all three GEMMs (forward, dgrad, wgrad) in split precision."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        R, K = x.shape
        N = w.shape[0]
        assert w.shape[1] == K and w.stride(1) == 1
        y = torch.empty((R, N), dtype=torch.float32, device=x.device)
        gemm3(R, N, K, x, K, 1, w, w.stride(0), 1, y, N, 1, bias=b)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        R, K = x.shape
        N = w.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)                                  # dx[r,k] = sum_n dy[r,n] w[n,k]
            gemm3(R, K, N, dy, N, 1, w, 1, w.stride(0), dx, K, 1)
        if ctx.needs_input_grad[1]:
            dw = torch.empty((N, K), dtype=torch.float32, device=x.device)   # dw[n,k] = sum_r dy[r,n] x[r,k]
            split = max(1, min(256, (R + 2047) // 2048))
            gemm3(N, K, R, dy, 1, N, x, 1, K, dw, K, 1, split_k=split)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db


def linear3(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    return _Linear3.apply(x, w, b)


def _rms_norm(x: torch.Tensor, scale: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    rms = x.norm(2, dim=-1, keepdim=True) / math.sqrt(x.shape[-1])
    return scale * (x / (rms + eps))


def _mlp_tail(mlp, h: torch.Tensor) -> torch.Tensor:
    """Layers 2..4 (+ RMSNorm) of a build_mlp container, given the first layer's pre-activation."""
    h = torch.relu(h)
    h = torch.relu(linear3(h, mlp[2].weight, mlp[2].bias))
    h = torch.relu(linear3(h, mlp[4].weight, mlp[4].bias))
    h = linear3(h, mlp[6].weight, mlp[6].bias)
    return _rms_norm(h, mlp[7].scale) if len(mlp) > 7 else h


def mlp3(mlp, x: torch.Tensor) -> torch.Tensor:
    return _mlp_tail(mlp, linear3(x, mlp[0].weight, mlp[0].bias))


def epd_forward(model, graph) -> torch.Tensor:
    """EncodeProcessDecode.forward (graphphysics/models/processors.py:162-215) in split precision."""
    x, ea, ei = graph.x.float(), graph.edge_attr.float(), graph.edge_index
    if x.device.type != "cuda":
        raise RuntimeError("graphphysics_b200 runs on CUDA devices only (there is no CPU fallback)")
    src, dst = ei[0], ei[1]
    H = model.hidden_size
    if not model.only_processor:
        x = mlp3(model.nodes_encoder, x)
        e = mlp3(model.edges_encoder, ea)
    else:
        e = ea
    for blk in model.processor_list:
        w1e, w1n = blk.edge_block[0].weight, blk.node_block[0].weight          # [H,3H] = [W1e|W1d|W1s], [H,2H] = [W1x|W1a]
        wp = torch.cat([w1e[:, H:2 * H], w1e[:, 2 * H:], w1n[:, :H]], 0)       # P = x . [W1d; W1s; W1x]^T
        P = linear3(x, wp, None)
        z1 = linear3(e, w1e[:, :H], blk.edge_block[0].bias) + P[dst, :H] + P[src, H:2 * H]
        u = _mlp_tail(blk.edge_block, z1)
        agg = torch.zeros_like(x).index_add_(0, dst, u)
        z1n = linear3(agg, w1n[:, H:], blk.node_block[0].bias) + P[:, 2 * H:]
        x = x + _mlp_tail(blk.node_block, z1n)
        e = e + u
    return x if model.only_processor else mlp3(model.decode_module, x)
