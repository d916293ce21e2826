"""TensorFolder (reference: utils/tensor_folder.py:6-98) — shape bookkeeping only."""
from typing import List, Sequence, Tuple

import torch


class TensorFolder:

    @staticmethod
    def prod(input: Sequence):
        result = 1
        for i in input:
            result *= i
        return result

    @staticmethod
    def flatten(tensor: torch.Tensor, dimensions: int = 2) -> Tuple[torch.Tensor, List]:
        size = list(tensor.size())
        if dimensions <= 0:
            dimensions = len(size) + dimensions
        return torch.reshape(tensor, tuple([TensorFolder.prod(size[:dimensions])] + size[dimensions:])), size[:dimensions]

    @staticmethod
    def flatten_list(tensors: List[torch.Tensor], dimensions: int = 2):
        first, dims = TensorFolder.flatten(tensors[0], dimensions)
        return [first] + [TensorFolder.flatten(t, dimensions)[0] for t in tensors[1:]], dims

    @staticmethod
    def fold(tensor: torch.Tensor, dimensions: List[int]) -> torch.Tensor:
        size = list(tensor.size())
        product = TensorFolder.prod(dimensions)
        if product != 0 and size[0] % product != 0:
            raise Exception(f"First dimension {size[0]} is not the product of the specified dimensions, nor dim1 can be inferred {dimensions}")
        if size[0] != product:
            dimensions = [size[0] // product] + list(dimensions)
        return torch.reshape(tensor, list(dimensions) + size[1:])

    @staticmethod
    def fold_list(tensors: List[torch.Tensor], dimensions: List[int]):
        return [TensorFolder.fold(t, dimensions) for t in tensors]
