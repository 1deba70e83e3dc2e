"""Masked L2 loss (graphphysics/utils/loss.py:19-75) on the CUDA kernel gp_masked_mse."""
from typing import Sequence

import torch
from torch.nn.modules.loss import _Loss

from .. import ops
from .nodetype import NodeType


def prepare_mask(node_type: torch.Tensor, masks: Sequence[int]) -> torch.Tensor:
    """uint8 row mask: 1 where node_type is one of `masks` (loss.py:19-34)."""
    m = node_type == int(masks[0])
    for t in masks[1:]:
        m = m | (node_type == int(t))
    return m.to(torch.uint8)


class _MaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target, mask_u8):
        out_c, tgt_c = out.contiguous().float(), target.contiguous().float()
        loss = torch.empty(1, dtype=torch.float32, device=out.device)
        grad = torch.empty_like(out_c)
        ops.masked_mse(out_c, tgt_c, mask_u8, loss, grad)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


class L2Loss(_Loss):
    """mean over the rows whose node type is in `masks` (x all output columns) of (out - target)^2."""

    @property
    def __name__(self):
        return "MSE"

    def forward(self, target, network_output, node_type, masks=(NodeType.NORMAL, NodeType.OUTFLOW), selected_indexes=None,
                **kwargs):
        mask = prepare_mask(node_type, masks)
        if selected_indexes is not None:
            keep = torch.ones_like(mask)
            keep[selected_indexes] = 0
            mask = mask & keep
        return _MaskedMSE.apply(network_output, target, mask)
