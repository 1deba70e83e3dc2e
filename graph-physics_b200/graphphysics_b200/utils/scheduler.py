"""Cosine learning-rate schedule with linear warm-up (graphphysics/utils/scheduler.py:8-67)."""
import math
from typing import List

import torch


def lr_factor(step_index: int, warmup: int, max_iters: int, min_lr_factor: float = 1e-3) -> float:
    """Factor applied to the base LR at scheduler index `step_index` (== `last_epoch`)."""
    t = step_index + 1
    f = 0.5 * (1.0 + math.cos(math.pi * t / max_iters))
    if t <= warmup:
        f *= t / warmup
    return max(f, min_lr_factor)


class CosineWarmupScheduler(torch.optim.lr_scheduler._LRScheduler):
    """Drop-in for the reference class: same constructor, same `get_lr_factor`."""

    def __init__(self, optimizer, warmup: int, max_iters: int, min_lr_factor: float = 0.001, last_epoch: int = -1):
        self.warmup, self.max_iters, self.min_lr_factor = warmup, max_iters, min_lr_factor
        super().__init__(optimizer, last_epoch)

    def get_lr_factor(self, epoch: int) -> float:
        return lr_factor(epoch, self.warmup, self.max_iters, self.min_lr_factor)

    def get_lr(self) -> List[float]:
        f = self.get_lr_factor(self.last_epoch)
        return [base * f for base in self.base_lrs]
