"""CPU train step of the reference algorithm, assembled from the oracle (TEST INFRASTRUCTURE:
used by tests, smoke() and bench.py's cpu_baseline / --impl reference legs only).

Restates LightningModule.training_step + configure_optimizers
(graphphysics/training/lightning_module.py:270-342, 494-511) and the Trainer knobs of
graphphysics/train.py:276-290 around the oracle's Simulator / EncodeProcessDecode / L2Loss:
fp32, torch autograd, clip_grad_norm_(1.0), AdamW(wd 1e-4, betas (0.9, 0.95)), cosine warm-up."""
from __future__ import annotations

from typing import Dict

import torch

from . import gp_oracle as O


def default_state_dict(num_layers: int, node_in: int, edge_in: int, out_size: int, hidden: int, seed: int = 0,
                       dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reference default init (nn.Linear kaiming-uniform, RMSNorm scale 1) under a fixed seed, with
    the reference's state_dict keys."""
    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def mlp(prefix, i, o, norm=True):
        for j, (n, k) in enumerate([(hidden, i), (hidden, hidden), (hidden, hidden), (o, hidden)]):
            lin = torch.nn.Linear(k, n)
            sd[f"{prefix}.{2 * j}.weight"] = lin.weight.detach().to(dtype)
            sd[f"{prefix}.{2 * j}.bias"] = lin.bias.detach().to(dtype)
        if norm:
            sd[f"{prefix}.7.scale"] = torch.ones(o, dtype=dtype)

    mlp("nodes_encoder", node_in, hidden)
    mlp("edges_encoder", edge_in, hidden)
    mlp("decode_module", hidden, out_size, norm=False)
    for l in range(num_layers):
        mlp(f"processor_list.{l}.edge_block", 3 * hidden, hidden)
        mlp(f"processor_list.{l}.node_block", 2 * hidden, hidden)
    return sd


class CpuTrainer:
    def __init__(self, sd: Dict[str, torch.Tensor], num_layers: int, index: Dict[str, int], out_size: int, node_in: int,
                 edge_in: int, lr: float, num_steps: int, warmup: int, mode=None):
        self.params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        self.L, self.index, self.mode = num_layers, index, mode
        dt = next(iter(sd.values())).dtype
        self.norms = {"output": O.Normalizer(out_size, dtype=dt), "node": O.Normalizer(node_in, dtype=dt),
                      "edge": O.Normalizer(edge_in, dtype=dt)}
        self.opt = torch.optim.AdamW(list(self.params.values()), lr=lr, weight_decay=1e-4, betas=(0.9, 0.95))
        self.base_lr, self.num_steps, self.warmup, self.step_index = lr, num_steps, warmup, 0

    def forward(self, x_raw, y, edge_attr, edge_index, training: bool):
        fn = lambda nf, ef: O.epd_forward(self.params, nf, ef, edge_index, self.L, mode=self.mode)
        return O.simulator_forward(fn, self.norms, x_raw, y, edge_attr, self.index, training)

    def training_step(self, x_raw, y, edge_attr, edge_index) -> float:
        out, target, _ = self.forward(x_raw, y, edge_attr, edge_index, True)
        loss = O.l2_loss(target, out, x_raw[:, self.index["node_type_index"]])
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(self.params.values()), 1.0)
        for gpar in self.opt.param_groups:
            gpar["lr"] = self.base_lr * O.cosine_warmup_factor(self.step_index, self.warmup, self.num_steps)
        self.opt.step()
        self.step_index += 1
        return float(loss.detach())
