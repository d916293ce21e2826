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

    # ------------------------------------------------------------------------------------------------------------------
    # ray selection (reference :55-183, 236-431, 583-795).  Same results as the reference for the same torch RNG state --
    # the random draws are made per image in the reference's order -- but tensorised over images and objects: no
    # ``.item()`` device->host sync and no Python loop over pixels, so a training step can stay asynchronous.
    # ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def coordinate_from_flat_index(index, image_width: int):
        """Reference :583-596."""
        return index // image_width, index % image_width

    @staticmethod
    def flat_index_from_coordinate(coordinate, image_width: int):
        """Reference :598-610."""
        row, column = coordinate
        return (row * image_width) + column

    @staticmethod
    def permutation_indices_to_positions(permutation_indices: torch.Tensor, height: int, width: int) -> torch.Tensor:
        """Reference :1157-1178: flat pixel indices -> (row / height, column / width)."""
        rows = permutation_indices // width
        cols = permutation_indices % width
        return torch.stack([rows / height, cols / width], dim=-1)

    @staticmethod
    def _hwc(observations: torch.Tensor) -> torch.Tensor:
        return observations.movedim(-3, -1)

    @staticmethod
    def _slice_bound(idx: torch.Tensor, size: int) -> torch.Tensor:
        """Python slice semantics of ``tensor[a:b]`` for an integer bound (negative values count from the end)."""
        return torch.where(idx < 0, (idx + size).clamp(min=0), idx.clamp(max=size))

    @staticmethod
    def bounding_box_weight_masks(bounding_boxes: torch.Tensor, weights: Sequence[float], height: int, width: int) -> torch.Tensor:
        """Spatial sampling weights of reference :97-123 / :313-338 / :647-674: every object adds
        ``weights[o] / area(o)`` inside its pixel-aligned (floor/ceil) box.  bounding_boxes (S, 4, objects) normalised
        (left, top, right, bottom) -> (S, height, width).  Boxes of zero area add nothing (the reference does the same in
        ``sample_rays_weighted`` and divides by zero in the patch samplers)."""
        device = bounding_boxes.device
        left = torch.floor(bounding_boxes[:, 0, :] * width).long()
        right = torch.ceil(bounding_boxes[:, 2, :] * width).long()
        top = torch.floor(bounding_boxes[:, 1, :] * height).long()
        bottom = torch.ceil(bounding_boxes[:, 3, :] * height).long()
        area = (right - left) * (bottom - top)
        ys = torch.arange(height, device=device).view(1, height, 1)
        xs = torch.arange(width, device=device).view(1, 1, width)
        masks = torch.zeros((bounding_boxes.size(0), height, width), dtype=torch.float32, device=device)
        for o in range(bounding_boxes.size(-1)):                     # objects in order: same float summation as the reference
            t0 = RayHelper._slice_bound(top[:, o], height).view(-1, 1, 1)
            b0 = RayHelper._slice_bound(bottom[:, o], height).view(-1, 1, 1)
            l0 = RayHelper._slice_bound(left[:, o], width).view(-1, 1, 1)
            r0 = RayHelper._slice_bound(right[:, o], width).view(-1, 1, 1)
            inside = (ys >= t0) & (ys < b0) & (xs >= l0) & (xs < r0)
            a = area[:, o].view(-1, 1, 1)
            # python float division, rounded to fp32 when added to the mask -- like the reference's ``+=`` of a python scalar
            value = (float(weights[o]) / a.double().clamp(min=1)).float()
            masks = masks + torch.where(inside & (a != 0), value, torch.zeros_like(value))
        return masks

    @staticmethod
    def _weighted_pixel_samples(masks: torch.Tensor, count: int, cdf_samples: torch.Tensor = None) -> torch.Tensor:
        """``count`` pixel indices per image drawn from the (S, H, W) weight masks by inverse-CDF sampling (reference
        :139-148).  The uniform numbers are drawn per image, in order, like the reference's ``torch.rand`` calls."""
        flat = masks.reshape(masks.size(0), -1)
        cdf = torch.cumsum(flat / flat.sum(dim=1, keepdim=True), dim=1)
        if cdf_samples is None:
            cdf_samples = torch.stack([torch.rand((count,), device=masks.device) for _ in range(masks.size(0))], dim=0)
        idx = torch.searchsorted(cdf, cdf_samples.contiguous())
        return torch.clamp(idx, max=cdf.size(1) - 1)

    @staticmethod
    def _gather_pixels(flat: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """flat (S, H*W, C), idx (S, n) (negative indices count from the end, like python indexing) -> (S, n, C)."""
        n_pix = flat.size(1)
        idx = torch.where(idx < 0, idx + n_pix, idx).long()
        return torch.gather(flat, 1, idx.unsqueeze(-1).expand(-1, -1, flat.size(-1)))

    @staticmethod
    def sample_rays(ray_directions: torch.Tensor, observations: torch.Tensor, samples_per_image: int):
        """Reference :730-795: ``samples_per_image`` uniformly chosen rays per image (0 = all rays, natural order)."""
        lead = list(ray_directions.shape[:-3])
        height, width = ray_directions.size(-3), ray_directions.size(-2)
        flat_d = ray_directions.reshape(-1, height * width, 3)
        flat_o = RayHelper._hwc(observations).reshape(-1, height * width, observations.size(-3))
        if samples_per_image > 0:
            perm = torch.stack([torch.randperm(height * width, device=ray_directions.device)[:samples_per_image]
                                for _ in range(flat_d.size(0))], dim=0)
            d, o = RayHelper._gather_pixels(flat_d, perm), RayHelper._gather_pixels(flat_o, perm)
        else:
            perm = torch.arange(height * width, dtype=ray_directions.dtype, device=ray_directions.device).repeat(flat_d.size(0), 1)
            d, o = flat_d, flat_o
        pos = RayHelper.permutation_indices_to_positions(perm, height, width)
        return TensorFolder.fold(d, lead), TensorFolder.fold(o, lead), TensorFolder.fold(pos, lead)

    @staticmethod
    def sample_rays_weighted(ray_directions: torch.Tensor, observations: torch.Tensor, samples_per_image: int,
                             bounding_boxes: torch.Tensor, weights: Sequence[float], cdf_samples: torch.Tensor = None):
        """Reference :611-728: rays drawn in proportion to the per-object bounding-box weights."""
        if samples_per_image <= 0:
            return RayHelper.sample_rays(ray_directions, observations, 0)
        lead = list(ray_directions.shape[:-3])
        height, width = ray_directions.size(-3), ray_directions.size(-2)
        flat_d = ray_directions.reshape(-1, height * width, 3)
        flat_o = RayHelper._hwc(observations).reshape(-1, height * width, observations.size(-3))
        boxes = bounding_boxes.reshape(-1, bounding_boxes.size(-2), bounding_boxes.size(-1))
        masks = RayHelper.bounding_box_weight_masks(boxes, weights, height, width)
        idx = RayHelper._weighted_pixel_samples(masks, samples_per_image, cdf_samples)
        pos = RayHelper.permutation_indices_to_positions(idx, height, width)
        return (TensorFolder.fold(RayHelper._gather_pixels(flat_d, idx), lead), TensorFolder.fold(RayHelper._gather_pixels(flat_o, idx), lead),
                TensorFolder.fold(pos, lead))

    @staticmethod
    def sample_rays_patched(ray_directions: torch.Tensor, observations: torch.Tensor, patch_size: int, patch_count: int,
                            bounding_boxes: torch.Tensor, weights: Sequence[float], cdf_samples: torch.Tensor = None):
        """Reference :55-183: ``patch_count`` dense patch_size x patch_size patches per image around weighted centres."""
        if patch_size % 2 != 0:
            raise Exception("Patch size must be a multiple of 2")
        lead = list(ray_directions.shape[:-3])
        height, width = ray_directions.size(-3), ray_directions.size(-2)
        flat_d = ray_directions.reshape(-1, height * width, 3)
        flat_o = RayHelper._hwc(observations).reshape(-1, height * width, observations.size(-3))
        boxes = bounding_boxes.reshape(-1, bounding_boxes.size(-2), bounding_boxes.size(-1))
        masks = RayHelper.bounding_box_weight_masks(boxes, weights, height, width)
        centre = RayHelper._weighted_pixel_samples(masks, patch_count, cdf_samples)              # (S, patches)
        half = patch_size // 2
        row = (centre // width).clamp(min=half).clamp(max=height - half) - half
        col = (centre % width).clamp(min=half).clamp(max=width - half) - half
        off = torch.arange(patch_size, device=centre.device)
        rows = row.unsqueeze(-1) + off                                                                # (S, patches, ps)
        cols = col.unsqueeze(-1) + off
        idx = (rows.unsqueeze(-1) * width + cols.unsqueeze(-2)).reshape(centre.size(0), -1)
        return TensorFolder.fold(RayHelper._gather_pixels(flat_d, idx), lead), TensorFolder.fold(RayHelper._gather_pixels(flat_o, idx), lead)

    @staticmethod
    def sample_rays_strided_patch(ray_directions: torch.Tensor, observations: torch.Tensor, patch_size: int, strides,
                                  bounding_boxes: torch.Tensor, weights: Sequence[float], align_grid=False,
                                  cdf_samples: torch.Tensor = None):
        """Reference :236-431: one multi-stride patch per image.  The patch centre is drawn from the bounding-box weights,
        moved inside the image and snapped so that every sampled ray is the centre pixel of a (stride x stride) cell; the
        samples of each stride (smallest first, ``patch_size * strides[0] / stride`` per side) are concatenated."""
        if not align_grid:
            raise Exception("Align grid is required for patched ray sampling.")
        if patch_size % 2 != 0:
            raise Exception("Patch size must be a multiple of 2")
        if not isinstance(strides, (list, tuple)):
            strides = [strides]
        smallest, biggest = strides[0], strides[-1]
        if (patch_size * smallest) % (2 * biggest) != 0:
            raise Exception("Patch size is not compatible with the chosen strides. Make patch size divisible by a higher power of 2")
        patch_sizes = [(patch_size * smallest) // s for s in strides]
        half = patch_sizes[-1] // 2                                   # half patch size at the biggest stride

        lead = list(ray_directions.shape[:-3])
        height, width = ray_directions.size(-3), ray_directions.size(-2)
        device = ray_directions.device
        flat_d = ray_directions.reshape(-1, height * width, 3)
        flat_o = RayHelper._hwc(observations).reshape(-1, height * width, observations.size(-3))
        boxes = bounding_boxes.reshape(-1, bounding_boxes.size(-2), bounding_boxes.size(-1))
        masks = RayHelper.bounding_box_weight_masks(boxes, weights, height, width)
        centre = RayHelper._weighted_pixel_samples(masks, 1, cdf_samples)[:, 0]                     # (S,)

        backward_map = torch.tensor(list(range(biggest // 2, biggest)) + list(range(0, biggest // 2)), device=device)
        forward_map = torch.tensor(list(range(biggest // 2 + biggest, biggest, -1)) + [0] + list(range(biggest - 1, biggest // 2, -1)),
                                   device=device)

        def start_of(c: torch.Tensor, size: int) -> torch.Tensor:
            c = c.clamp(min=half * biggest).clamp(max=size - biggest * (half - 1) - 1)            # keep the patch inside the image
            start = c - half * biggest
            diff = start % biggest                                                                 # snap to the cell centres
            moved = torch.where(start >= biggest // 2, start - backward_map[diff], start + forward_map[diff])
            return torch.where(diff != biggest // 2, moved, start)

        start_row, start_col = start_of(centre // width, height), start_of(centre % width, width)
        all_idx = []
        for s, ps in zip(strides, patch_sizes):
            offset = biggest // 2 - s // 2
            steps = torch.arange(ps, device=device) * s
            rows = (start_row - offset).unsqueeze(-1) + steps                                      # (S, ps)
            cols = (start_col - offset).unsqueeze(-1) + steps
            all_idx.append((rows.unsqueeze(-1) * width + cols.unsqueeze(-2)).reshape(centre.size(0), -1))
        idx = torch.cat(all_idx, dim=1)
        pos = RayHelper.permutation_indices_to_positions(idx.int(), height, width)
        return (TensorFolder.fold(RayHelper._gather_pixels(flat_d, idx), lead), TensorFolder.fold(RayHelper._gather_pixels(flat_o, idx), lead),
                TensorFolder.fold(pos, lead))

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
    def ray_parameters(z_near: torch.Tensor, z_far: torch.Tensor, positions_count: int, rand: torch.Tensor = None) -> torch.Tensor:
        """The ray parameters ``t`` of create_ray_positions (reference :1253-1277) for per-ray bounds (..., R): uniform in [z_near, z_far],
        jittered inside their strata by ``rand`` (..., R, P) when given."""
        s = torch.linspace(0.0, 1.0, positions_count, device=z_near.device)
        t = z_near.unsqueeze(-1) * (1.0 - s) + z_far.unsqueeze(-1) * s
        if rand is not None:
            mid = (t[..., 1:] + t[..., :-1]) / 2
            upper = torch.cat([mid, t[..., -1:]], dim=-1)
            lower = torch.cat([t[..., :1], mid], dim=-1)
            t = lower + (upper - lower) * rand
        return t

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

    # ------------------------------------------------------------------------------------------------------------------
    # hierarchical ("fine") sampling helpers (reference :1284-1403), used by ObjectComposer's fine pass (``use_fine: True``; no shipped
    # config enables it).
    # ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def transform_ray_positions(ray_origins, ray_directions, focal_normals, ray_positions, transformation_matrix):
        """Reference :1284-1318 (broadcast instead of the per-sample matrix ``repeat``)."""
        o, d, n = RayHelper.transform_rays(ray_origins, ray_directions, focal_normals, transformation_matrix)
        p = RayHelper.transform_points(ray_positions, transformation_matrix.unsqueeze(-3).unsqueeze(-3))
        return o, d, n, p

    @staticmethod
    def sample_pdf(bin_delimiters: torch.Tensor, weights: torch.Tensor, positions_count: int, perturb: bool,
                   cdf_samples: torch.Tensor = None) -> torch.Tensor:
        """Reference :1349-1403: inverse-CDF samples of a piecewise-constant density over the bins between ``bin_delimiters``
        (..., B) with ``weights`` (..., B - 1).  (The reference adds its 1e-5 to ``weights`` in place; this version does not
        modify its argument.)"""
        weights = weights + 1e-5
        pdf = weights / torch.sum(weights, dim=-1, keepdim=True)
        cdf = torch.cumsum(pdf, dim=-1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
        lead = list(cdf.shape[:-1])
        if cdf_samples is None:
            if not perturb:
                cdf_samples = torch.linspace(0.0, 1.0, positions_count, device=cdf.device).expand(lead + [positions_count])
            else:
                cdf_samples = torch.rand(lead + [positions_count]).to(cdf.device)
        cdf_samples = cdf_samples.contiguous()
        idx = torch.searchsorted(cdf, cdf_samples, right=True)
        below = torch.clamp(idx - 1, min=0)
        above = torch.clamp(idx, max=cdf.size(-1) - 1)
        cdf_lo, cdf_hi = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
        bin_lo, bin_hi = torch.gather(bin_delimiters, -1, below), torch.gather(bin_delimiters, -1, above)
        norm = cdf_hi - cdf_lo
        norm = torch.where(norm < 1e-5, torch.ones_like(norm), norm)
        return bin_lo + (cdf_samples - cdf_lo) / norm * (bin_hi - bin_lo)

    @staticmethod
    def create_ray_positions_weighted(ray_origins, ray_directions, positions_count: int, reference_ray_positions_t, weights,
                                      perturb: bool):
        """Reference :1320-1347: ``positions_count`` new samples drawn from the coarse weights, merged (sorted) with the coarse ones."""
        mid = (reference_ray_positions_t[..., 1:] + reference_ray_positions_t[..., :-1]) / 2
        t_new = RayHelper.sample_pdf(mid, weights[..., 1:-1], positions_count, perturb).detach()
        merged, _ = torch.sort(torch.cat([reference_ray_positions_t, t_new], dim=-1), dim=-1)
        positions = ray_origins.unsqueeze(-2).unsqueeze(-2) + ray_directions.unsqueeze(-2) * merged.unsqueeze(-1)
        return positions, merged

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
