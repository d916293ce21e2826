"""PositionalRayBender (reference: model/nerf_models/positional_ray_bender_model.py:12-175).

annealed PE(x/size) || deformation -> ``layers_count`` x [Linear(width)+ReLU] (input re-concatenated at
``skip_layer_idx``) -> Linear(3, no bias) * size, clamped so that x + displacement stays inside the box.
Parameter container + initialisation; evaluated in the field kernel in front of the NeRF field."""
from typing import Dict

import torch
import torch.nn as nn

from ..annealable_positional_encoder import AnnealablePositionalEncoder
from ...utils.lib_3d.bounding_box import BoundingBox
from ... import _cabi


class PositionalRayBender(nn.Module):
    KIND = _cabi.BENDER_POSITIONAL

    def __init__(self, config: Dict, model_config: Dict):
        super().__init__()
        self.config = config
        self.model_config = model_config
        self.layers_width = model_config["layers_width"]
        self.layers_count = model_config["layers_count"]
        self.skip_layer_idx = model_config["skip_layer_idx"]
        self.deformation_features = model_config["deformation_features"]
        pe = model_config["position_encoder"]
        if not pe["append_original"]:
            raise Exception("the B200 field kernels implement append_original=True (every shipped config)")
        self.positional_encoder = AnnealablePositionalEncoder(3, pe["octaves"], pe["append_original"], pe["num_steps"])
        self.bounding_box = BoundingBox(model_config["bounding_box"])
        self.last_layer_bias = False
        self.backbone_layers = nn.ModuleList()
        current = self.positional_encoder.get_encoding_size() + self.deformation_features
        for layer_idx in range(self.layers_count):
            if layer_idx == self.skip_layer_idx:
                current += self.positional_encoder.get_encoding_size() + self.deformation_features
            self.backbone_layers.append(nn.Linear(current, self.layers_width))
            current = self.layers_width
        self.output_head = nn.Linear(self.layers_width, 3, bias=self.last_layer_bias)
        self.init_weights()

    def set_step(self, current_step: int):
        self.positional_encoder.set_step(current_step)

    def init_weights(self):
        """Reference :66-79 (including its quirk: the near-zero init lands on the LAST BACKBONE layer)."""
        for layer in self.backbone_layers:
            torch.nn.init.kaiming_uniform_(layer.weight, a=0, mode="fan_in", nonlinearity="relu")
            torch.nn.init.zeros_(layer.bias)
        torch.nn.init.uniform_(self.backbone_layers[-1].weight, a=-1e-5, b=1e-5)

    def forward(self, ray_positions: torch.Tensor, deformation: torch.Tensor, video_indexes: torch.Tensor = None):
        raise NotImplementedError("the ray bender is evaluated inside RayBendingStyleNerfModel (fused field kernel); "
                                  "call the parent model, its third output is the displacement")


def model(config, model_config):
    return PositionalRayBender(config, model_config)
