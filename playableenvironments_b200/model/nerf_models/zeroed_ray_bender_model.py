"""ZeroedRayBender (reference: model/nerf_models/zeroed_ray_bender_model.py:7-49): never bends rays."""
from typing import Dict

import torch
import torch.nn as nn

from ... import _cabi


class ZeroedRayBender(nn.Module):
    KIND = _cabi.BENDER_ZEROED

    def __init__(self, config: Dict, model_config: Dict):
        super().__init__()
        self.config = config
        self.model_config = model_config

    def set_step(self, current_step: int):
        pass

    def forward(self, ray_positions: torch.Tensor, deformation: torch.Tensor, video_indexes: torch.Tensor = None):
        return ray_positions * 0.0


def model(config, model_config):
    return ZeroedRayBender(config, model_config)
