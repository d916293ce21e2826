"""AdaInStyleNerfModel (reference: model/nerf_models/adain_style_nerf_model.py:14-209).

PE(3) -> ``backbone_layers_count`` x [Linear(width)+ReLU] with the encoding re-concatenated at ``skip_layer_idx`` ->
alpha head (1) || feature head [Linear -> AdaIn -> ReLU -> Linear -> AdaIn -> ReLU -> Linear(output_features)].
This class owns the parameters under the reference's state_dict names; evaluation happens in the fused field kernels
(csrc/pe_field_tc.cu for the shipped 8x256/192 shape, csrc/pe_field_fp32.cu otherwise)."""
from typing import Dict, Tuple

import torch
import torch.nn as nn

from ..layers.adain import AffineTransformAdaIn
from ..layers.adain_sequential import AdaInSequential
from ..positional_encoder import PositionalEncoder
from ...utils.lib_3d.bounding_box import BoundingBox
from ... import _cabi


class AdaInStyleNerfModel(nn.Module):
    KIND = _cabi.NERF_ADAIN
    INPUT_DIMENSIONS = 3
    HAS_ALPHA_HEAD = True

    def __init__(self, config: Dict, model_config: Dict):
        super().__init__()
        self.config = config
        self.model_config = model_config
        self.layers_width = model_config["layers_width"]
        self.backbone_layers_count = model_config["backbone_layers_count"]
        self.output_features = model_config["output_features"]
        self.skip_layer_idx = model_config["skip_layer_idx"]
        self.style_features = model_config["style_features"]
        self.empty_space_alpha = model_config["empty_space_alpha"]
        if self.skip_layer_idx >= self.backbone_layers_count:
            raise Exception("Skip layer must refer to a valid backbone layer idx")
        self.position_encoder = PositionalEncoder(self.INPUT_DIMENSIONS, model_config["position_encoder"]["octaves"],
                                                  model_config["position_encoder"]["append_original"])
        if not model_config["position_encoder"]["append_original"]:
            raise Exception("the B200 field kernels implement append_original=True (every shipped config)")
        self.bounding_box = BoundingBox(model_config["bounding_box"])
        self.backbone_layers = nn.ModuleList()
        current = self.position_encoder.get_encoding_size()
        for layer_idx in range(self.backbone_layers_count):
            if layer_idx == self.skip_layer_idx:
                current += self.position_encoder.get_encoding_size()
            self.backbone_layers.append(nn.Linear(current, self.layers_width))
            current = self.layers_width
        if self.HAS_ALPHA_HEAD:
            self.alpha_head = nn.Linear(self.layers_width, 1)
        self.features_head = self.get_features_head()

    def get_features_head(self):
        cls = self.get_style_embedding_layer_class()
        return AdaInSequential(
            nn.Linear(self.layers_width, self.layers_width, bias=False),
            cls(self.layers_width, self.style_features),
            nn.ReLU(),
            nn.Linear(self.layers_width, self.layers_width // 2, bias=False),
            cls(self.layers_width // 2, self.style_features),
            nn.ReLU(),
            nn.Linear(self.layers_width // 2, self.output_features))

    def get_style_embedding_layer_class(self):
        return AffineTransformAdaIn

    def compute_bounding_box_filtering_mask(self, flat_ray_positions: torch.Tensor) -> torch.Tensor:
        return self.bounding_box.is_inside(flat_ray_positions)

    def forward(self, ray_positions: torch.Tensor, ray_origins: torch.Tensor, ray_directions: torch.Tensor, style: torch.Tensor,
                video_indexes: torch.Tensor = None) -> Tuple[torch.Tensor]:
        """(..., 3) positions, (..., S) style -> (..., F) features, (...) raw alphas, {} — reference :147-199.
        Evaluated by wrapping the field in a bender-free object and calling the field kernel on explicit positions."""
        from .ray_bending_style_nerf_model import evaluate_field_on_positions
        feats, alphas, _ = evaluate_field_on_positions(self, None, ray_positions, ray_origins, ray_directions, style, None,
                                                       self.training, False)
        return feats, alphas, {}


def model(config, model_config):
    return AdaInStyleNerfModel(config, model_config)
