"""TensorBroadcaster (reference: utils/tensor_broadcaster.py:4-28).  Returns an expanded VIEW: the fused kernels never
need the materialised per-sample copies the reference creates with ``repeat``."""
import torch


class TensorBroadcaster:

    @staticmethod
    def add_dimension(tensor: torch.Tensor, size: int, dim: int) -> torch.Tensor:
        tensor = tensor.unsqueeze(dim)
        sizes = [-1] * tensor.dim()
        sizes[dim] = size
        return tensor.expand(sizes)
