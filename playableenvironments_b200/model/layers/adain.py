"""AffineTransformAdaIn / AdaIn (reference: model/layers/adain.py:5-61).

Parameter containers with the reference's state_dict names (``affine_transform.{weight,bias}``,
``ada_in.normalization.{running_mean,running_var,num_batches_tracked}``).  The arithmetic —
``BatchNorm1d(affine=False)(x) * scale + bias`` with ``[scale|bias] = Linear(style)`` — runs inside the field kernels
(style prologue + layer epilogue); ``forward`` here only serves callers that use the layer on its own."""
import torch
import torch.nn as nn


class AdaIn(nn.Module):

    def __init__(self, in_features: int):
        super().__init__()
        self.normalization = nn.BatchNorm1d(in_features, affine=False)

    def forward(self, input, scale, bias):
        return self.normalization(input) * scale + bias


class AffineTransformAdaIn(nn.Module):

    def __init__(self, in_features: int, style_features_count: int):
        super().__init__()
        self.style_features_count = style_features_count
        self.affine_transform = nn.Linear(self.style_features_count, 2 * in_features)
        self.ada_in = AdaIn(in_features)
        self.affine_transform.bias.data[:in_features] = 1     # scale biased to 1, bias to 0 (reference :17-19)
        self.affine_transform.bias.data[in_features:] = 0

    def forward(self, input: torch.Tensor, style: torch.Tensor):
        scale, bias = self.affine_transform(style).chunk(2, 1)
        return self.ada_in(input, scale, bias)
