"""BoundingBox (reference: utils/lib_3d/bounding_box.py:7-131): axis-aligned box buffer; the kernels take its 6 floats."""
from typing import Sequence

import torch
import torch.nn as nn


class BoundingBox(nn.Module):

    def __init__(self, dimensions: Sequence):
        super().__init__()
        if len(dimensions) != 3:
            raise Exception(f"Dimenions should have dimension 3, but dimension ({len(dimensions)}) was passed")
        self.register_buffer("dimensions", torch.as_tensor(dimensions, dtype=torch.float32), persistent=False)
        self._flat = [float(v) for pair in dimensions for v in pair]       # host copy: no device read on the hot path

    def as_floats(self):
        return list(self._flat)

    def get_center_offset(self, device=None) -> torch.Tensor:
        center = self.dimensions[:, 0] + (self.dimensions[:, 1] - self.dimensions[:, 0]) / 2
        return center if device is None else center.to(device)

    def is_inside(self, points: torch.Tensor):
        return torch.logical_and(torch.all(points <= self.dimensions[:, 1], dim=-1), torch.all(points >= self.dimensions[:, 0], dim=-1))

    def get_size(self) -> torch.Tensor:
        return self.dimensions[:, 1] - self.dimensions[:, 0]

    def get_corner_points(self) -> torch.Tensor:
        """(8, 3); point 0 = all low, point 6 = all high (reference numbering, bounding_box.py:58-98)."""
        lo, hi = self.dimensions[:, 0], self.dimensions[:, 1]
        pick = [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1), (0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1)]
        rows = [torch.stack([hi[a] if sel[a] else lo[a] for a in range(3)]) for sel in pick]
        return torch.stack(rows)

    def get_edge_points(self, points_per_edge: int = 5) -> torch.Tensor:
        idx = [0, 1, 1, 2, 2, 3, 3, 0, 4, 5, 5, 6, 6, 7, 7, 4, 0, 4, 1, 5, 2, 6, 3, 7]
        corners = self.get_corner_points()
        ends = corners[idx].reshape(12, 2, 3)
        frac = torch.linspace(0.0, 1.0, points_per_edge + 2, device=corners.device)[1:-1]
        pts = ends[:, 0].unsqueeze(-1) + (ends[:, 1] - ends[:, 0]).unsqueeze(-1) * frac
        return torch.cat([corners, pts.transpose(1, 2).reshape(-1, 3)], dim=0)
