"""Host-side launch of the fused render path through the C ABI (``pe_render_forward``).

Everything here is plumbing: flatten the leading (B, O, C) dims into ``images``, hand raw device pointers of torch
tensors to the library on the caller's current stream, wrap the outputs into the reference's nested result dict.
No arithmetic of the render path happens in Python."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from .. import _cabi

INTEGRATED_KEYS = ("integrated_features", "opacity", "weights", "depth", "disparity",
                   "integrated_displacements_magnitude", "integrated_divergence")

_launches = 0       # kernels launched by render calls in this process (bench.py reports it)


def take_launch_count() -> int:
    return int(_cabi.lib().pe_take_launch_count())


class _Workspace:
    """One cached workspace per device and stream: pe_render_forward never allocates."""
    _cache: Dict = {}

    @classmethod
    def get(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            cls._cache[key] = buf
        return buf


def _alloc_integrated(images_shape: Sequence[int], rays: int, P: int, F: int, device) -> Dict[str, torch.Tensor]:
    lead = list(images_shape) + [rays]
    e = lambda *s: torch.empty(lead + list(s), dtype=torch.float32, device=device)
    return {"integrated_features": e(F), "opacity": e(), "weights": e(P), "depth": e(), "disparity": e(),
            "integrated_displacements_magnitude": e(), "integrated_divergence": e()}


def _fill(struct: _cabi.PeIntegrated, tensors: Dict[str, torch.Tensor]):
    for k in INTEGRATED_KEYS:
        setattr(struct, k, _cabi.ptr(tensors[k]))


def render_scene(descs: List[_cabi.PeObjectDesc], static_objects: int, ray_origins: torch.Tensor, ray_directions: torch.Tensor,
                 w2o: torch.Tensor, style: torch.Tensor, deformation: torch.Tensor, object_in_scene: torch.Tensor,
                 perturb: bool, training: bool, fix_object_overlaps: bool, apply_activation: bool, precision: int,
                 rand: Optional[List[torch.Tensor]] = None, noise: Optional[Dict[str, torch.Tensor]] = None,
                 bn_running: Optional[List] = None, return_raw_alphas: bool = False) -> Dict:
    """One ObjectComposer.forward (reference: model/object_composer.py:786-892).  Returns {"object_k": {...}, "global": {...}}."""
    device = ray_directions.device
    if device.type != "cuda":
        raise _cabi.PeError("ObjectComposer.forward needs CUDA tensors: the B200 render path has no CPU implementation")
    L = _cabi.lib()
    lead = list(ray_directions.shape[:-2])
    rays = ray_directions.size(-2)
    images = 1
    for v in lead:
        images *= v
    K = len(descs)
    if K > _cabi.PE_MAX_OBJECTS:
        raise _cabi.PeError(f"at most {_cabi.PE_MAX_OBJECTS} object instances per composer call")
    F = descs[0].features

    scene = _cabi.PeScene()
    scene.images, scene.rays, scene.objects, scene.static_objects = images, rays, K, static_objects
    scene.perturb, scene.training = int(bool(perturb)), int(bool(training))
    scene.fix_object_overlaps, scene.apply_activation = int(bool(fix_object_overlaps)), int(bool(apply_activation))
    scene.precision, scene.explicit_positions = precision, 0
    for k, d in enumerate(descs):
        scene.object[k] = d

    keep = []

    def dev(t, shape):
        c = _cabi.f32(t.expand(shape) if list(t.shape) != list(shape) else t).reshape(-1)
        keep.append(c)
        return c

    ins = _cabi.PeInputs()
    ins.ray_origins = _cabi.ptr(dev(ray_origins, lead + [3]))
    ins.ray_directions = _cabi.ptr(dev(ray_directions, lead + [rays, 3]))
    # (..., 4, 4, objects) -> [images][objects][3][4]
    m = w2o.expand(lead + [4, 4, K]).reshape(images, 4, 4, K)[:, :3, :, :].permute(0, 3, 1, 2)
    ins.w2o = _cabi.ptr(dev(m, [images, K, 3, 4]))
    S, D = style.size(-2), deformation.size(-2)
    sty = style.expand(lead + [S, style.size(-1)]).reshape(images, S, -1)
    dfm = deformation.expand(lead + [D, deformation.size(-1)]).reshape(images, D, -1)
    for k in range(K):
        ins.style[k] = _cabi.ptr(dev(sty[:, :, k], [images, S]))
        ins.deformation[k] = _cabi.ptr(dev(dfm[:, :, k], [images, D]))
    ois = object_in_scene.expand(lead + [object_in_scene.size(-1)]).reshape(images, -1)[:, :K].to(torch.uint8).contiguous()
    keep.append(ois)
    ins.object_in_scene = _cabi.ptr(ois)
    if perturb:
        if rand is None or noise is None:
            # torch's generator replaces the reference's torch.rand (ray_helper.py:1275) / torch.randn (object_composer.py:194)
            rand = [torch.rand(lead + [rays, d.positions], device=device) for d in descs]
            noise = {f"object_{k}": torch.randn(lead + [rays, d.positions], device=device) for k, d in enumerate(descs)}
            noise["global"] = torch.randn(lead + [rays, sum(d.positions for d in descs)], device=device)
        for k, d in enumerate(descs):
            ins.rand[k] = _cabi.ptr(dev(rand[k], lead + [rays, d.positions]))
            ins.noise[k] = _cabi.ptr(dev(noise[f"object_{k}"], lead + [rays, d.positions]))
        ins.noise_global = _cabi.ptr(dev(noise["global"], lead + [rays, sum(d.positions for d in descs)]))

    outs = _cabi.PeOutputs()
    results: Dict = {}
    for k, d in enumerate(descs):
        r = _alloc_integrated(lead, rays, d.positions, F, device)
        _fill(outs.object[k], r)
        results[f"object_{k}"] = r
        if return_raw_alphas:          # diagnostic: per-sample raw alphas (first return value family of the object models)
            r["raw_alphas"] = torch.empty(lead + [rays, d.positions], dtype=torch.float32, device=device)
            outs.raw_alphas[k] = _cabi.ptr(r["raw_alphas"])
        if training and bn_running is not None:
            b1 = torch.empty((2, d.width), dtype=torch.float32, device=device)
            b2 = torch.empty((2, d.width // 2), dtype=torch.float32, device=device)
            outs.bn1_running[k], outs.bn2_running[k] = _cabi.ptr(b1), _cabi.ptr(b2)
            bn_running.append((b1, b2))
    g = _alloc_integrated(lead, rays, sum(d.positions for d in descs), F, device)
    _fill(outs.global_, g)
    results["global"] = g

    with torch.cuda.device(device):
        nbytes = L.pe_workspace_bytes(C.byref(scene))
        if nbytes == 0:
            raise _cabi.PeError(f"pe_workspace_bytes: {L.pe_last_error().decode()}")
        ws = _Workspace.get(device, nbytes)
        _cabi.check(L.pe_render_forward(C.byref(scene), C.byref(ins), C.byref(outs), _cabi.ptr(ws), ws.numel(),
                                        _cabi.current_stream(device)))
    del keep
    return results


def field_on_positions(desc: _cabi.PeObjectDesc, images: int, n: int, positions, origins, directions, style, deformation,
                       training: bool, device):
    """RayBendingStyleNerfModel.forward on explicit positions (reference: ray_bending_style_nerf_model.py:137-219)."""
    if device.type != "cuda":
        raise _cabi.PeError("the field operator needs CUDA tensors: there is no CPU implementation")
    L = _cabi.lib()
    scene = _cabi.PeScene()
    scene.images, scene.rays, scene.objects, scene.static_objects = images, n, 1, 1
    scene.training = int(bool(training))
    scene.precision, scene.explicit_positions = _cabi.PRECISION_FP32, 1
    scene.object[0] = desc
    ins = _cabi.PeInputs()
    ins.positions, ins.ray_origins, ins.ray_directions = _cabi.ptr(positions), _cabi.ptr(origins), _cabi.ptr(directions)
    ins.style[0] = _cabi.ptr(style)
    ins.deformation[0] = _cabi.ptr(deformation)
    feats = torch.empty((images, n, desc.features), dtype=torch.float32, device=device)
    alphas = torch.empty((images, n), dtype=torch.float32, device=device)
    disp = torch.empty((images, n, 3), dtype=torch.float32, device=device)
    outs = _cabi.PeOutputs()
    outs.raw_features[0], outs.raw_alphas[0], outs.displacements[0] = _cabi.ptr(feats), _cabi.ptr(alphas), _cabi.ptr(disp)
    with torch.cuda.device(device):
        nbytes = L.pe_workspace_bytes(C.byref(scene))
        ws = _Workspace.get(device, nbytes)
        _cabi.check(L.pe_render_forward(C.byref(scene), C.byref(ins), C.byref(outs), _cabi.ptr(ws), ws.numel(),
                                        _cabi.current_stream(device)))
    return feats, alphas, disp
