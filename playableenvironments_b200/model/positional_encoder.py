"""PositionalEncoder (reference: model/positional_encoder.py:4-65): Fourier features
``[x, sin(2^k x), cos(2^k x)]_k``, no pi factor.  Inside the render path the encoding is fused into the field
kernels; ``forward`` is the stand-alone operator (``pe_positional_encoding``)."""
import torch
import torch.nn as nn

from .. import _cabi


class PositionalEncoder(nn.Module):

    def __init__(self, input_dimensions: int, octaves_count: int, append_original: bool):
        super().__init__()
        self.input_dimensions = input_dimensions
        self.octaves_count = octaves_count
        self.append_original = append_original
        self.register_buffer("octaves", 2.0 ** torch.linspace(0.0, octaves_count - 1, octaves_count), persistent=False)

    def get_encoding_size(self) -> int:
        size = 2 * self.octaves_count * self.input_dimensions
        if self.append_original:
            size += self.input_dimensions
        return size

    def _weights(self):
        return None

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if input.size(-1) != self.input_dimensions:
            raise Exception(f"Input dimension ({input.size(-1)}) differs from expected input dimension ({self.input_dimensions})")
        x = _cabi.f32(input).reshape(-1, self.input_dimensions)
        out = torch.empty((x.size(0), self.get_encoding_size()), dtype=torch.float32, device=x.device)
        w = self._weights()
        w = None if w is None else _cabi.f32(w.to(x.device))
        _cabi.check(_cabi.lib().pe_positional_encoding(_cabi.ptr(x), x.size(0), self.input_dimensions, self.octaves_count,
                                                       1 if self.append_original else 0, _cabi.ptr(w), _cabi.ptr(out),
                                                       _cabi.current_stream(x.device)))
        return out.reshape(list(input.shape[:-1]) + [self.get_encoding_size()])
