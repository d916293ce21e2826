"""AnnealablePositionalEncoder (reference: model/annealable_positional_encoder.py:8-76): per-octave cosine-eased weight
driven by the ``current_step`` buffer (kept in the state_dict under the same name)."""
import math

import torch

from .positional_encoder import PositionalEncoder


def annealing_weights(current_step: int, octaves_count: int, num_steps: int):
    """(1 - cos(pi * clamp(step * octaves / num_steps - k, 0, 1))) / 2 — reference :54-58, evaluated in fp32 like torch."""
    alpha = torch.tensor(float(current_step), dtype=torch.float32) * octaves_count / num_steps
    idx = torch.arange(octaves_count, dtype=torch.float32)
    return (1 - torch.cos(math.pi * torch.clamp(alpha - idx, min=0.0, max=1.0))) / 2


class AnnealablePositionalEncoder(PositionalEncoder):

    def __init__(self, input_dimensions: int, octaves_count: int, append_original: bool, num_steps: int):
        super().__init__(input_dimensions, octaves_count, append_original)
        self.num_steps = num_steps
        self.register_buffer("current_step", torch.zeros((), dtype=torch.int))
        self.register_buffer("octave_indexes", torch.arange(octaves_count, dtype=torch.float32), persistent=False)
        self._host_step = 0            # host mirror: the kernel launch must not read device memory

    def set_step(self, current_step: int):
        self._host_step = int(current_step)
        self.current_step = self.current_step * 0 + current_step

    def host_step(self) -> int:
        return self._host_step

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "current_step"
        if key in state_dict:
            self._host_step = int(state_dict[key])
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _weights(self):
        return annealing_weights(self._host_step, self.octaves_count, self.num_steps)
