"""RayHelper — the ray helpers on the render path (reference: utils/lib_3d/ray_helper.py).

Inside ``ObjectComposer.forward`` the transforms, slab test and sample placement are fused into the field kernels, so the
per-sample tensors these helpers create in the reference never exist.  The static methods below keep the reference's
signatures for callers that use them directly; they are thin tensor-shape code or stand-alone kernels
(``generate_strided_grid_rays`` = create_camera_rays + sample_all_rays_strided_grid + transform_rays in one launch)."""
import ctypes as C
from typing import List, Sequence, Tuple, Union

import torch

from ... import _cabi
from ..tensor_folder import TensorFolder


class RayHelper:

    @staticmethod
    def create_camera_rays(initial_dimensions: List[int], height: int, width: int, focal, device=None):
        """Reference :15-52.  Pinhole rays in the camera frame (not normalised, pixel centre at the integer index)."""
        if not torch.is_tensor(focal):
            focal = torch.full(list(initial_dimensions), float(focal), dtype=torch.float32, device=device or "cuda")
        device = focal.device
        focal = focal.unsqueeze(-1).unsqueeze(-1)
        rows, cols = torch.meshgrid(torch.arange(0, height, device=device), torch.arange(0, width, device=device), indexing="ij")
        dx = (cols - width / 2) / focal
        dy = -(rows - height / 2) / focal
        dz = -torch.ones_like(dx)
        directions = torch.stack([dx, dy, dz], -1)
        normals = torch.zeros(list(initial_dimensions) + [3], device=device)
        normals[..., 2] = -1
        return directions, torch.zeros_like(normals), normals

    @staticmethod
    def generate_strided_grid_rays(focal: torch.Tensor, camera_to_world: torch.Tensor, height: int, width: int,
                                   strides: Union[Sequence[int], int]):
        """Fused create_camera_rays (:15-52) + sample_all_rays_strided_grid (:433-482) + transform_rays (:1203-1227).

        focal (...,), camera_to_world (..., 4, 4) -> world-space directions (..., R, 3), origins (..., 3),
        normalised (row, col) positions (..., R, 2) with R = sum_s (H/s)(W/s).  The full (..., H, W, 3) ray grid of the
        reference is never materialised."""
        if not isinstance(strides, (list, tuple)):
            strides = [strides]
        lead = list(focal.shape)
        images = 1
        for v in lead:
            images *= v
        R = sum((height // s) * (width // s) for s in strides)
        device = focal.device
        f = _cabi.f32(focal).reshape(-1)
        m = _cabi.f32(camera_to_world.expand(lead + [4, 4]).reshape(images, 4, 4)[:, :3, :])
        directions = torch.empty((images, R, 3), dtype=torch.float32, device=device)
        origins = torch.empty((images, 3), dtype=torch.float32, device=device)
        positions = torch.empty((images, R, 2), dtype=torch.float32, device=device)
        arr = (C.c_int32 * len(strides))(*strides)
        with torch.cuda.device(device):
            _cabi.check(_cabi.lib().pe_generate_rays(_cabi.ptr(f), _cabi.ptr(m), images, height, width, arr, len(strides),
                                                     _cabi.ptr(directions), _cabi.ptr(origins), _cabi.ptr(positions),
                                                     _cabi.current_stream(device)))
        return directions.reshape(lead + [R, 3]), origins.reshape(lead + [3]), positions.reshape(lead + [R, 2])

    @staticmethod
    def sample_strided_grid(tensor: torch.Tensor, stride: int):
        """Reference :533-582: centre pixel ``idx*stride + stride//2`` of every stride x stride cell."""
        h, w = tensor.size(-3), tensor.size(-2)
        if h % stride != 0:
            raise Exception("The image height is not divisible by the stride")
        if w % stride != 0:
            raise Exception("The image width is not divisible by the stride")
        off = stride // 2
        rows = torch.arange(h // stride, device=tensor.device) * stride + off
        cols = torch.arange(w // stride, device=tensor.device) * stride + off
        out = tensor.index_select(-3, rows).index_select(-2, cols)
        idx = torch.stack(torch.meshgrid(rows.float() / h, cols.float() / w, indexing="ij"), dim=-1)
        idx = idx.expand(list(out.shape[:-3]) + list(idx.shape))
        return out, idx

    @staticmethod
    def sample_all_rays_strided_grid(ray_directions: torch.Tensor, observations: torch.Tensor, strides):
        """Reference :433-482."""
        if not isinstance(strides, (list, tuple)):
            strides = [strides]
        observations = observations.movedim(-3, -1)
        all_d, all_i, all_o = [], [], []
        for s in strides:
            d, i = RayHelper.sample_strided_grid(ray_directions, s)
            o, _ = RayHelper.sample_strided_grid(observations, s)
            all_d.append(d.reshape(list(d.shape[:-3]) + [-1, d.size(-1)]))
            all_i.append(i.reshape(list(i.shape[:-3]) + [-1, i.size(-1)]))
            all_o.append(o.reshape(list(o.shape[:-3]) + [-1, o.size(-1)]))
        return torch.cat(all_d, dim=-2), torch.cat(all_o, dim=-2), torch.cat(all_i, dim=-2)

    @staticmethod
    def fold_strided_grid_samples(samples: torch.Tensor, strides, original_size: Tuple[int], dim: int) -> List[torch.Tensor]:
        """Reference :484-531 (views only)."""
        if not isinstance(strides, (list, tuple)):
            strides = [strides]
        image_height, image_width = original_size
        out, start = [], 0
        for s in strides:
            if image_height % s != 0:
                raise Exception("The image height is not divisible by the stride")
            if image_width % s != 0:
                raise Exception("The image width is not divisible by the stride")
            gh, gw = image_height // s, image_width // s
            sl = [slice(None)] * samples.dim()
            sl[dim] = slice(start, start + gh * gw)
            cur = samples[tuple(sl)]
            shape = list(cur.shape)
            shape[dim:dim + 1] = [gh, gw]
            out.append(cur.reshape(shape))
            start += gh * gw
        return out

    @staticmethod
    def fold_feature_grids(features: torch.Tensor, strides: Sequence[int], original_size: Tuple[int, int],
                           channels: Sequence[int]) -> List[torch.Tensor]:
        """Decoder hand-off in one pass: fold_strided_tensors (environment_model_backpropagated_autoencoder.py:129-168) +
        per-stride channel split + HWC->CHW (environment_model_multiresolution_backpropagated_autoencoder.py:59-99).
        features (..., R, F) -> [(..., channels_i, H/s_i, W/s_i)]."""
        lead = list(features.shape[:-2])
        images = 1
        for v in lead:
            images *= v
        H, W = original_size
        F = features.size(-1)
        device = features.device
        feats = _cabi.f32(features).reshape(images, -1, F)
        grids = [torch.empty((images, c, H // s, W // s), dtype=torch.float32, device=device) for s, c in zip(strides, channels)]
        s_arr = (C.c_int32 * len(strides))(*strides)
        c_arr = (C.c_int32 * len(strides))(*channels)
        g_arr = (C.c_void_p * len(strides))(*[_cabi.ptr(g) for g in grids])
        with torch.cuda.device(device):
            _cabi.check(_cabi.lib().pe_fold_feature_grids(_cabi.ptr(feats), images, H, W, F, s_arr, c_arr, len(strides), g_arr,
                                                          _cabi.current_stream(device)))
        return [g.reshape(lead + list(g.shape[1:])) for g in grids]

    @staticmethod
    def transform_points(points: torch.Tensor, transformation_matrix: torch.Tensor, rotation=True, translation=True):
        """Reference :1180-1201."""
        out = points
        if rotation:
            out = torch.sum(out.unsqueeze(-2) * transformation_matrix[..., :3, :3], -1)
        if translation:
            out = out + transformation_matrix[..., :3, -1]
        return out

    @staticmethod
    def transform_rays(ray_origins, ray_directions, focal_normals, transformation_matrix):
        """Reference :1203-1227 (broadcast instead of the reference's per-ray matrix ``repeat``)."""
        o = RayHelper.transform_points(ray_origins, transformation_matrix)
        n = RayHelper.transform_points(focal_normals, transformation_matrix, translation=False)
        d = RayHelper.transform_points(ray_directions, transformation_matrix.unsqueeze(-3), translation=False)
        return o, d, n

    @staticmethod
    def create_ray_positions(ray_origins, ray_directions, z_near, z_far, positions_count: int, perturb: bool):
        """Reference :1229-1282."""
        device = ray_directions.device
        if not torch.is_tensor(z_near):
            z_near = torch.ones(ray_origins.size()[:-1], dtype=torch.float32, device=device) * z_near
        if not torch.is_tensor(z_far):
            z_far = torch.ones(ray_origins.size()[:-1], dtype=torch.float32, device=device) * z_far
        s = torch.linspace(0.0, 1.0, positions_count, device=device)
        t = z_near.unsqueeze(-1) * (1.0 - s) + z_far.unsqueeze(-1) * s
        if z_near.dim() == ray_origins.dim() - 1:
            t = t.unsqueeze(-2).expand(list(ray_directions.shape[:-1]) + [positions_count])
        if perturb:
            mid = (t[..., 1:] + t[..., :-1]) / 2
            upper = torch.cat([mid, t[..., -1:]], dim=-1)
            lower = torch.cat([t[..., :1], mid], dim=-1)
            t = lower + (upper - lower) * torch.rand(t.size(), device=device)
        positions = ray_origins.unsqueeze(-2).unsqueeze(-2) + ray_directions.unsqueeze(-2) * t.unsqueeze(-1)
        return positions, t

    @staticmethod
    def strided_patch_ray_samples_to_patch(samples: torch.Tensor):
        """Reference :185-204."""
        samples = torch.transpose(samples, -1, -2)
        patch = int(round(samples.size(-1) ** 0.5))
        return samples.reshape(list(samples.shape[:-1]) + [patch, patch])

    @staticmethod
    def split_strided_patch_ray_samples(samples: torch.Tensor, patch_size: int, strides) -> List[torch.Tensor]:
        """Reference :206-234."""
        if not isinstance(strides, (list, tuple)):
            strides = [strides]
        out, begin = [], 0
        for s in strides:
            size = (patch_size * strides[0]) // s
            out.append(samples[..., begin:begin + size ** 2, :])
            begin += size ** 2
        return out
