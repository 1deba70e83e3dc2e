"""Simulator wrapper (graphphysics/models/simulator.py:13-275): feature assembly, online
normalisation, target delta and de-normalised outputs around the processor.  Plain PyTorch
elementwise work on [N, few] tensors; the model call underneath is the CUDA path."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..graph import Data
from ..utils.nodetype import NodeType
from .layers import Normalizer


class Simulator(nn.Module):
    def __init__(self, node_input_size: int, edge_input_size: int, output_size: int, feature_index_start: int,
                 feature_index_end: int, output_index_start: int, output_index_end: int, node_type_index: int,
                 model: nn.Module, device: torch.device, model_dir: str = "checkpoint/simulator.pth"):
        super().__init__()
        self.node_input_size = node_input_size
        self.edge_input_size = edge_input_size if edge_input_size > 0 else None
        self.output_size = output_size
        self.feature_index_start, self.feature_index_end = feature_index_start, feature_index_end
        self.output_index_start, self.output_index_end = output_index_start, output_index_end
        self.node_type_index = node_type_index
        self.model_dir = model_dir
        self.model = model.to(device)
        self._output_normalizer = Normalizer(size=output_size, name="output_normalizer", device=device)
        self._node_normalizer = Normalizer(size=node_input_size, name="node_normalizer", device=device)
        self._edge_normalizer = (Normalizer(size=edge_input_size, name="edge_normalizer", device=device)
                                 if self.edge_input_size is not None else None)
        self.device = device

    # simulator.py:80-110
    def _get_pre_target(self, inputs) -> torch.Tensor:
        return inputs.x[:, self.output_index_start:self.output_index_end]

    def _get_target_normalized(self, inputs, is_training: bool = True) -> torch.Tensor:
        return self._output_normalizer(inputs.y - self._get_pre_target(inputs), is_training)

    # simulator.py:112-143
    def _get_one_hot_type(self, inputs) -> torch.Tensor:
        return F.one_hot(inputs.x[:, self.node_type_index].long(), NodeType.SIZE)

    def _build_node_features(self, inputs, one_hot_type: torch.Tensor) -> torch.Tensor:
        return torch.cat([inputs.x[:, self.feature_index_start:self.feature_index_end], one_hot_type], dim=1)

    def node_features(self, inputs) -> torch.Tensor:
        """[features | one-hot node type] (simulator.py:112-143): one kernel for CUDA fp32 inputs (gp_node_features)."""
        x = inputs.x
        if x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1:
            from .. import ops
            return ops.node_features(x, self.feature_index_start, self.feature_index_end, self.node_type_index, NodeType.SIZE)
        return self._build_node_features(inputs, self._get_one_hot_type(inputs)).float()

    # simulator.py:145-176
    def _build_input_graph(self, inputs, is_training: bool):
        target_delta_normalized = self._get_target_normalized(inputs, is_training)
        node_features = self._node_normalizer(self.node_features(inputs), is_training)
        edge_attr = inputs.edge_attr
        if self._edge_normalizer is not None:
            edge_attr = self._edge_normalizer(edge_attr, is_training)
        graph = Data(x=node_features, pos=inputs.pos, edge_attr=edge_attr, edge_index=inputs.edge_index)
        return graph, target_delta_normalized

    # simulator.py:178-191
    def build_outputs(self, inputs, network_output: torch.Tensor) -> torch.Tensor:
        return self._get_pre_target(inputs) + self._output_normalizer.inverse(network_output)

    # simulator.py:193-217
    def forward(self, inputs) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        graph, target_delta_normalized = self._build_input_graph(inputs, self.training)
        network_output = self.model(graph)
        if self.training:
            return network_output, target_delta_normalized, None
        return network_output, target_delta_normalized, self.build_outputs(inputs, network_output)

    def freeze_all(self) -> None:
        for p in self.model.parameters():
            p.requires_grad = False

    # simulator.py:226-275 -- same file layout, so reference checkpoints load here and vice versa
    def load_checkpoint(self, ckpdir: Optional[str] = None) -> None:
        ckpt = torch.load(ckpdir or self.model_dir, map_location=self.device)
        self.load_state_dict(ckpt["model"])
        for key in ("_output_normalizer", "_node_normalizer", "_edge_normalizer"):
            state, norm = ckpt.get(key) or {}, getattr(self, key, None)
            if norm is not None:
                for attr in ("_acc_count", "_num_accumulations", "_acc_sum", "_acc_sum_squared"):
                    if attr in state:
                        getattr(norm, attr).copy_(state[attr])
                norm._host_calls = None

    def save_checkpoint(self, savedir: Optional[str] = None) -> None:
        savedir = savedir or self.model_dir
        os.makedirs(os.path.dirname(savedir) or ".", exist_ok=True)
        torch.save({"model": self.state_dict(),
                    "_output_normalizer": self._output_normalizer.get_variable(),
                    "_node_normalizer": self._node_normalizer.get_variable(),
                    "_edge_normalizer": self._edge_normalizer.get_variable() if self._edge_normalizer else None},
                   savedir)
