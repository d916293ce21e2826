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
    BatchNorm statistics would be per rank; training shards whole frames per rank instead (``allreduce_gradients`` below)."""
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


class PeerGather:
    """The all-gather of the rendered feature grid FUSED into the render kernel (the one collective of the path).

    Every rank owns ``buffer`` (world, rays, F) in symmetric memory (CUDA virtual-memory allocations exchanged once at construction:
    torch.distributed._symmetric_memory) and maps the buffers of the other ranks of the box into its own address space.  Handing
    ``destinations()`` to ``ObjectComposer.forward(peer_features=...)`` makes the field kernel store each ray's features into slot
    ``rank`` of EVERY rank's buffer as it produces them: P2P stores over NVLink / NVSwitch, no NCCL kernel competing for the SMs (the
    persistent render kernel leaves none free) and nothing exposed after the render but ``sync()``, a device-side barrier on the
    symmetric memory's signal pads.  (Legacy cudaIpc mappings fault under kernel stores from the importing device; measured.)
    Inference only; one image per call."""

    def __init__(self, rays: int, features: int, device: torch.device, group=None):
        self.group, self.rays, self.features = group, rays, features
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        shape = (self.world, rays, features)
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            self.buffer = symm_mem.empty(shape, dtype=torch.float32, device=device)
            self._handle = symm_mem.rendezvous(self.buffer, group=group if group is not None else dist.group.WORLD)
            self._peers = [self.buffer if r == self.rank else self._handle.get_buffer(r, shape, torch.float32) for r in range(self.world)]
            self.buffer.zero_()
            torch.cuda.synchronize(device)
            dist.barrier(group=group)
        else:
            self.buffer = torch.zeros(shape, dtype=torch.float32, device=device)
            self._handle, self._peers = None, [self.buffer]

    def destinations(self, begin: int = 0, end: int = None) -> List[torch.Tensor]:
        """Slot ``rank`` (rays ``begin:end``) of every rank's buffer, the local one included."""
        end = self.rays if end is None else end
        return [self._peers[r][self.rank, begin:end] for r in range(self.world)]

    def sync(self):
        """Orders every rank's reads of ``buffer`` after every rank's render (stream-ordered, on the current stream)."""
        if self._handle is not None:
            self._handle.barrier()
        return self.buffer


def render_pipelined(composer, ray_origins, ray_directions, focal_normals, w2o, style, deformation, object_in_scene, perturb: bool,
                     chunks: int = 4, group=None, gathered: torch.Tensor = None, host_out: torch.Tensor = None, side_stream=None,
                     peer_gather: "PeerGather" = None, **kw):
    """Renders the local rays in ``chunks`` contiguous ray chunks and moves each chunk's feature grid off the compute stream while
    the next chunk renders: chunk i is all-gathered across the ranks (``gathered``: (world, rays, F), every rank renders the same
    number of rays -- one frame or one equal shard per rank) and/or copied to pinned host memory (``host_out``: (rays, F)) on
    ``side_stream``; only the last chunk's transfer is exposed.  With ``peer_gather`` the all-gather is fused into the render kernel
    instead (P2P stores, ``PeerGather``) and only the host copy rides the side stream.  Rays are independent, so the chunks equal the
    single-launch render bit for bit.  Inference only.  Returns the local feature grid (rays, F) (device)."""
    if torch.is_grad_enabled() and composer.training:
        raise Exception("render_pipelined is inference-only: call it under torch.no_grad() / composer.eval()")
    world = dist.get_world_size(group) if (dist.is_initialized() and gathered is not None) else 1
    rays = ray_directions.size(-2)
    if ray_directions.numel() != rays * 3:
        raise Exception("render_pipelined renders ONE image per call (leading dims of size 1): ray chunks of several images are not "
                        "contiguous in the gathered grid; call it per image")
    main = torch.cuda.current_stream()
    side = side_stream if side_stream is not None else torch.cuda.Stream()
    local, keep = None, []
    for c in range(chunks):
        begin, end = ray_shard(rays, c, chunks)
        if end == begin:
            continue
        if peer_gather is not None:
            kw["peer_features"] = peer_gather.destinations(begin, end)
        res = composer(ray_origins, ray_directions[..., begin:end, :], focal_normals, w2o, style, deformation, object_in_scene, perturb, **kw)
        feats = res["coarse"]["global"]["integrated_features"]
        feats = feats.reshape(end - begin, feats.size(-1))
        if peer_gather is not None and host_out is None:
            local = peer_gather.buffer[peer_gather.rank]
            continue
        if local is None:
            local = feats.new_empty((rays, feats.size(-1))) if (gathered is None or world == 1) else None
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            if gathered is not None and world > 1:
                dist.all_gather([gathered[r, begin:end] for r in range(world)], feats, group=group)
            elif gathered is not None:
                gathered[0, begin:end].copy_(feats, non_blocking=True)
            if local is not None:
                local[begin:end].copy_(feats, non_blocking=True)
            if host_out is not None:
                host_out[begin:end].copy_(feats, non_blocking=True)
        keep.append(feats)                  # alive until the side stream has consumed it
    done = torch.cuda.Event()
    done.record(side)
    main.wait_event(done)
    for t in keep:
        t.record_stream(side)
    if peer_gather is not None:
        peer_gather.sync()
        return peer_gather.buffer[peer_gather.rank]
    if local is None:
        local = gathered[dist.get_rank(group)]
    return local


def allreduce_gradients(parameters, group=None, average: bool = True) -> int:
    """Data-parallel training across the ranks (the reference: nn.DataParallel's replica-gradient reduction, train.py:61): every
    rank renders whole frames, then ONE all-reduce of a single flat bucket holding every composer gradient (2.87 M parameters =
    11.5 MB in the shipped Tennis configuration) -- sized for launch latency, not for link count (NVSwitch).  Parameters without a
    gradient on this rank contribute zeros.  Train-mode BatchNorm statistics stay per rank, like the reference's replicas.
    Returns the number of bytes reduced."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    if world > 1:
        dist.all_reduce(flat, group=group)
        if average:
            flat /= world
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * flat.element_size()
