"""JSON config -> model / simulator (graphphysics/training/parse_parameters.py:81-190).
The `model`, `index` and `training` sections of training_config/*.json are read verbatim."""
from __future__ import annotations

from typing import Any, Dict

import torch

from ..models.layers import set_memory_optimized_training, set_use_silu_activation
from ..models.processors import EncodeProcessDecode
from ..models.simulator import Simulator
from ..utils.nodetype import NodeType


def get_model(param: Dict[str, Any], only_processor: bool = False):
    m = param.get("model", {})
    model_type = m.get("type", "")
    node_input_size = param["model"]["node_input_size"] + NodeType.SIZE     # parse_parameters.py:96
    training = param.get("training", {})
    set_use_silu_activation(bool(m.get("use_silu_activation", False)))
    set_memory_optimized_training(training.get("enable_vram_optimizations", False))
    common = dict(use_rope_embeddings=m.get("use_rope_embeddings", False),
                  use_gated_attention=m.get("use_gated_attention", False),
                  rope_pos_dimension=m.get("rope_pos_dimension", 3), rope_base=m.get("rope_base", 10000.0),
                  use_temporal_block=training.get("use_temporal_block", False))
    if model_type == "epd":
        return EncodeProcessDecode(message_passing_num=m["message_passing_num"], node_input_size=node_input_size,
                                   edge_input_size=m["edge_input_size"], output_size=m["output_size"],
                                   hidden_size=m["hidden_size"], only_processor=only_processor,
                                   use_gated_mlp=m.get("use_gated_mlp", False), precision=m.get("precision"), **common)
    if model_type == "transformer":
        from ..models.processors import EncodeTransformDecode
        return EncodeTransformDecode(message_passing_num=m["message_passing_num"], node_input_size=node_input_size,
                                     output_size=m["output_size"], hidden_size=m["hidden_size"],
                                     num_heads=m["num_heads"], only_processor=only_processor, precision=m.get("precision"),
                                     **common)
    if model_type == "transolver":
        raise NotImplementedError("model type 'transolver' is outside the accelerated path (SURVEY §2)")
    raise ValueError(f"Model type '{model_type}' not supported.")


def get_simulator(param: Dict[str, Any], model, device: torch.device) -> Simulator:
    idx = param["index"]
    return Simulator(node_input_size=param["model"]["node_input_size"] + NodeType.SIZE,
                     edge_input_size=param["model"]["edge_input_size"], output_size=param["model"]["output_size"],
                     feature_index_start=idx["feature_index_start"], feature_index_end=idx["feature_index_end"],
                     output_index_start=idx["output_index_start"], output_index_end=idx["output_index_end"],
                     node_type_index=idx["node_type_index"], model=model, device=device)


def get_preprocessing(param: Dict[str, Any], device: torch.device = None, use_edge_feature: bool = True, remove_noise: bool = False,
                      extra_node_features=None, extra_edge_features=None):
    """transformations.preprocessing / transformations.world_pos_parameters of the JSON config -> the device-side transform
    pipeline (graphphysics/training/parse_parameters.py:24-78; the transforms are graphphysics_b200/preprocessing.py).
    `device` is accepted for call compatibility: the transforms run where the graph's tensors live (a CUDA device)."""
    from ..preprocessing import build_preprocessing
    pre = param.get("transformations", {}).get("preprocessing", {})
    noise_scale = pre.get("noise", 0)
    noise_parameters = None
    if noise_scale != 0 and not remove_noise:
        noise_parameters = {"noise_index_start": pre.get("noise_index_start"), "noise_index_end": pre.get("noise_index_end"),
                            "noise_scale": noise_scale, "node_type_index": param["index"]["node_type_index"]}
    wp = param.get("transformations", {}).get("world_pos_parameters", {})
    world_pos_parameters = None
    if wp.get("use", False):
        world_pos_parameters = {"world_pos_index_start": wp.get("world_pos_index_start"),
                                "world_pos_index_end": wp.get("world_pos_index_end"),
                                "node_type_index": param["index"]["node_type_index"]}
    return build_preprocessing(noise_parameters=noise_parameters, world_pos_parameters=world_pos_parameters,
                               add_edges_features=use_edge_feature, extra_node_features=extra_node_features,
                               extra_edge_features=extra_edge_features)


def get_loss(param: Dict[str, Any], **kwargs):
    """parse_parameters.py:300-323 for the loss of the path: (L2Loss(), "L2LOSS").  The physics-informed losses of
    graphphysics/utils/loss.py (divergence, gradient ...) and MultiLoss are outside the accelerated path."""
    from ..utils.loss import L2Loss
    types = [t.upper() for t in param.get("loss", {}).get("type", ["l2loss"])]
    if types != ["L2LOSS"]:
        raise NotImplementedError(f"loss types {types}: only L2Loss is part of the accelerated path (SURVEY §2)")
    return L2Loss(**kwargs), "L2LOSS"


def get_gradient_method(param: Dict[str, Any], **kwargs):
    """parse_parameters.py:326-341: loss.gradient_method of the config, or None."""
    return param.get("loss", {}).get("gradient_method")
