"""TensorBatchifier (reference: utils/tensor_batchifier.py:6-45).

Kept for API compatibility.  The fused render path keeps no per-sample activations in HBM, so
``batchified_composer_call`` (see model/environment_model_glue.py) renders the whole ray set in one call."""
from typing import List

import torch


class TensorBatchifier:

    @staticmethod
    def batchify(tensor: torch.Tensor, dim: int, batch_size: int) -> List[torch.Tensor]:
        if dim < 0:
            dim += tensor.dim()
        return list(torch.split(tensor, batch_size, dim=dim))
