"""Training step of the UNMODIFIED reference on the host CPU (TEST INFRASTRUCTURE: bench.py's `--impl reference` and
`cpu_baseline` legs only).

The model, simulator, loss and LR scheduler are the reference's own classes, imported from the copy staged by
oracle/build_ref.py (oracle/_ref/graphphysics, verified against its MANIFEST) through oracle/ref_shim.py.  Only the
loop Lightning runs around them is restated, from graphphysics/training/lightning_module.py:270-342 (training_step),
494-511 (configure_optimizers: AdamW(lr, weight_decay=1e-4, betas=(0.9, 0.95)) + CosineWarmupScheduler stepped every
iteration) and graphphysics/train.py:276-290 (gradient_clip_val=1.0)."""
from __future__ import annotations

import hashlib
import json
import os
from typing import Any, Dict

import torch

from . import ref_shim

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def verify_staged() -> str:
    man = os.path.join(_STAGED, "MANIFEST.json")
    if not os.path.exists(man):
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    for rel, digest in json.load(open(man))["files"].items():
        got = hashlib.sha256(open(os.path.join(_STAGED, rel), "rb").read()).hexdigest()
        if got != digest:
            raise RuntimeError(f"oracle/_ref/{rel} differs from the staged reference file")
    return _STAGED


class ReferenceTrainer:
    def __init__(self, config: Dict[str, Any], lr: float, num_steps: int, warmup: int, seed: int = 0):
        ref = ref_shim.import_reference(verify_staged())
        from torch_geometric.data import Data                      # the shim's attribute bag
        self.Data = Data
        m, index = config["model"], config["index"]
        torch.manual_seed(seed)
        # parse_parameters.py:96, 177: node input = JSON value + 9 (one-hot node type)
        self.net = ref["processors"].EncodeProcessDecode(m["message_passing_num"], m["node_input_size"] + 9,
                                                         m["edge_input_size"], m["output_size"], hidden_size=m["hidden_size"])
        self.sim = ref["simulator"].Simulator(node_input_size=m["node_input_size"] + 9, edge_input_size=m["edge_input_size"],
                                              output_size=m["output_size"], model=self.net, device=torch.device("cpu"), **index)
        self.loss = ref["loss"].L2Loss()
        NT = ref["nodetype"].NodeType
        self.masks = [NT.NORMAL, NT.OUTFLOW]
        self.opt = torch.optim.AdamW(self.sim.parameters(), lr=lr, weight_decay=1e-4, betas=(0.9, 0.95))
        self.sched = ref["scheduler"].CosineWarmupScheduler(self.opt, warmup=warmup, max_iters=num_steps)
        self.node_type_index = index["node_type_index"]
        self.sim.train()

    def training_step(self, x, y, pos, edge_index, edge_attr) -> float:
        b = self.Data(x=x, y=y, pos=pos, edge_index=edge_index, edge_attr=edge_attr)
        net, tgt, _ = self.sim(b)
        loss = self.loss(tgt, net, b.x[:, self.node_type_index], masks=self.masks)
        self.opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.sim.parameters(), 1.0)
        self.opt.step()
        self.sched.step()
        return float(loss.detach())
