"""SkyboxAdaInStyleNerfModelV3 (reference: model/nerf_models/skybox_adain_style_nerf_model_v3.py:14-159): the same trunk
and AdaIn head on PE(origin/size || unit direction) (6-D), alpha forced to 10.0, no bounding-box mask of its own."""
from typing import Dict

from ... import _cabi
from .adain_style_nerf_model import AdaInStyleNerfModel


class SkyboxAdaInStyleNerfModelV3(AdaInStyleNerfModel):
    KIND = _cabi.NERF_SKYBOX_V3
    INPUT_DIMENSIONS = 6
    HAS_ALPHA_HEAD = False

    def __init__(self, config: Dict, model_config: Dict):
        super().__init__(config, model_config)
        self.occupied_space_alpha = 10.0


def model(config, model_config):
    return SkyboxAdaInStyleNerfModelV3(config, model_config)
