"""Ray sharding across the GPUs of one box.

Rays are independent in the render path, so the ray dimension is split into contiguous chunks, one per rank, each rank renders
its chunk with no data-path communication, and ONE all-gather of the rendered feature grid (``integrated_features``) follows
before the shared decoder — the only collective on the path.  (The reference's nn.DataParallel — train.py:61 — scatters the
batch dim and gathers the whole nested result dict to GPU 0 instead.)"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def ray_shard(rays: int, rank: int, world_size: int, multiple: int = 1) -> Tuple[int, int]:
    """[begin, end) of the rays rendered by ``rank``: contiguous, sizes differ by at most ``multiple``."""
    units = (rays + multiple - 1) // multiple
    base, extra = divmod(units, world_size)
    begin_u = rank * base + min(rank, extra)
    end_u = begin_u + base + (1 if rank < extra else 0)
    return min(begin_u * multiple, rays), min(end_u * multiple, rays)


def shard_sizes(rays: int, world_size: int, multiple: int = 1) -> List[int]:
    return [ray_shard(rays, r, world_size, multiple)[1] - ray_shard(rays, r, world_size, multiple)[0] for r in range(world_size)]


def all_gather_rays(local: torch.Tensor, rays: int, dim: int, group=None, multiple: int = 1) -> torch.Tensor:
    """All-gathers a tensor sharded by ``ray_shard`` along ``dim`` back to the full ray set (one collective)."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = shard_sizes(rays, world, multiple)
    dim = dim % local.dim()
    moved = local.movedim(dim, 0).contiguous()
    pad = max(sizes)
    if moved.size(0) < pad:     # equal-sized buffers so a single all_gather_into_tensor moves everything
        moved = torch.cat([moved, moved.new_zeros((pad - moved.size(0),) + tuple(moved.shape[1:]))], dim=0)
    out = moved.new_empty((world * pad,) + tuple(moved.shape[1:]))
    dist.all_gather_into_tensor(out, moved, group=group)
    parts = [out[r * pad:r * pad + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0).movedim(0, dim)


def render_sharded(composer, ray_origins, ray_directions, focal_normals, w2o, style, deformation, object_in_scene, perturb: bool,
                   group=None, **kw):
    """Renders this rank's ray shard and all-gathers the global feature grid.  Returns
    (full integrated_features (..., R, F), local results dict, (begin, end)).

    INFERENCE ONLY: the all-gather is not autograd-aware (no gradient would reach the other ranks' shards) and train-mode
    BatchNorm statistics would be per rank; training shards whole frames per rank instead (``sharding.allreduce_gradients``)."""
    if torch.is_grad_enabled() and composer.training:
        raise Exception("render_sharded is inference-only: call it under torch.no_grad() / composer.eval()")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rays = ray_directions.size(-2)
    begin, end = ray_shard(rays, rank, world)
    local = composer(ray_origins, ray_directions[..., begin:end, :], focal_normals, w2o, style, deformation, object_in_scene, perturb, **kw)
    feats = local["coarse"]["global"]["integrated_features"]
    full = all_gather_rays(feats, rays, dim=-2, group=group) if world > 1 else feats
    return full, local, (begin, end)
