"""AdaInSequential (reference: model/layers/adain_sequential.py:10-28): Sequential that also feeds ``style`` to the
AdaIn members.  Container for the feature head ``features_head.{0,1,3,4,6}``."""
import torch
import torch.nn as nn

from .adain import AffineTransformAdaIn


class AdaInSequential(nn.Sequential):

    def forward(self, x: torch.Tensor, style: torch.Tensor):
        for module in self._modules.values():
            x = module(x, style) if isinstance(module, AffineTransformAdaIn) else module(x)
        return x
