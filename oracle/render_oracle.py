"""CPU oracle for the per-frame NeRF render path of PlayableEnvironments.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``playableenvironments_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it,
and there only as the checker / the CPU arm that is timed beside the GPU path.

This is a plain functional restatement (torch CPU ops on explicit tensors, no
``nn.Module``) of the reference algorithm.  Every function cites the reference
file:line it follows (paths relative to the upstream repository root).

Parity status: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the pins were generated here by importing the reference
itself: ``tests/golden/make_golden.py`` runs ``model.object_composer.ObjectComposer``
from the upstream tree on the seeded scenes of ``tests/golden/scenes.py`` and commits
its outputs under ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
this restatement against those files.

Parameters are passed as a flat ``dict[str, Tensor]`` using the reference's
``state_dict`` key names relative to one object model, e.g.
``nerf_model.backbone_layers.0.weight``.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5          # torch.nn.BatchNorm1d default, model/layers/adain.py:47
BN_MOMENTUM = 0.1      # torch.nn.BatchNorm1d default


# ----------------------------------------------------------------------------
# Geometry (utils/lib_3d/ray_helper.py, model/object_composer.py)
# ----------------------------------------------------------------------------

def create_camera_rays(initial_dimensions: Sequence[int], height: int, width: int, focal) -> Tuple[Tensor, Tensor, Tensor]:
    """utils/lib_3d/ray_helper.py:15-52.  Pinhole rays in camera frame, not
    normalised, pixel centre at the integer index, camera looks down -z."""
    if not torch.is_tensor(focal):
        focal = torch.full(list(initial_dimensions), float(focal), dtype=torch.float32)
    focal = focal.unsqueeze(-1).unsqueeze(-1)
    rows, cols = torch.meshgrid(torch.arange(0, height), torch.arange(0, width), indexing="ij")
    dx = (cols - width / 2) / focal
    dy = -(rows - height / 2) / focal
    dz = -torch.ones_like(dx)
    directions = torch.stack([dx, dy, dz], -1)
    normals = torch.zeros(list(initial_dimensions) + [3])
    normals[..., 2] = -1
    origins = torch.zeros_like(normals)
    return directions, origins, normals


def sample_strided_grid_directions(directions: Tensor, stride: int) -> Tuple[Tensor, Tensor]:
    """utils/lib_3d/ray_helper.py:533-582: centre pixel ``idx*stride + stride//2``
    of every stride x stride cell; positions normalised as row/H, col/W."""
    height, width = directions.size(-3), directions.size(-2)
    if height % stride or width % stride:
        raise Exception("The image size is not divisible by the stride")
    off = stride // 2
    rows = [i * stride + off for i in range(height // stride)]
    cols = [i * stride + off for i in range(width // stride)]
    out = directions[..., rows, :, :][..., cols, :]
    idx = torch.tensor([[[r / height, c / width] for c in cols] for r in rows], dtype=torch.float32)
    idx = idx.expand(list(out.shape[:-3]) + list(idx.shape)).contiguous()
    return out, idx


def sample_all_rays_strided_grid(directions: Tensor, strides) -> Tuple[Tensor, Tensor]:
    """utils/lib_3d/ray_helper.py:433-482 (directions and positions only)."""
    if not isinstance(strides, (list, tuple)):
        strides = [strides]
    all_d, all_p = [], []
    for s in strides:
        d, p = sample_strided_grid_directions(directions, s)
        all_d.append(d.reshape(list(d.shape[:-3]) + [-1, d.size(-1)]))
        all_p.append(p.reshape(list(p.shape[:-3]) + [-1, p.size(-1)]))
    return torch.cat(all_d, dim=-2), torch.cat(all_p, dim=-2)


def transform_points(points: Tensor, matrix: Tensor, rotation: bool = True, translation: bool = True) -> Tensor:
    """utils/lib_3d/ray_helper.py:1180-1201."""
    out = points
    if rotation:
        out = torch.sum(out.unsqueeze(-2) * matrix[..., :3, :3], -1)
    if translation:
        out = out + matrix[..., :3, -1]
    return out


def transform_rays(origins: Tensor, directions: Tensor, normals: Tensor, matrix: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """utils/lib_3d/ray_helper.py:1203-1227.  Directions are rotated only."""
    t_origins = transform_points(origins, matrix)
    t_normals = transform_points(normals, matrix, translation=False)
    t_directions = transform_points(directions, matrix.unsqueeze(-3), translation=False)
    return t_origins, t_directions, t_normals


def raywise_object_z_bounds(origins: Tensor, directions: Tensor, bbox: Tensor, object_validity: Tensor) -> Tuple[Tensor, Tensor]:
    """model/object_composer.py:104-151.  Slab test against the AABB with the
    reference's sign-unaware ``+1e-6`` on the direction; rays that miss, or
    objects absent from the scene, get near = far = 0."""
    eps = 1e-6
    corners = torch.stack([bbox[:, 0], bbox[:, 1]], dim=0)            # corner 0 (all low) and 6 (all high)
    corners = corners - origins.unsqueeze(-2)                          # (..., 2, 3)
    corners = corners.unsqueeze(-3)                                    # (..., 1, 2, 3)
    z = corners / (directions.unsqueeze(-2) + eps)                     # (..., R, 2, 3)
    z_near = z.min(dim=-2)[0].max(dim=-1)[0]
    z_far = z.max(dim=-2)[0].min(dim=-1)[0]
    validity = object_validity.unsqueeze(-1).expand_as(z_far)
    mask = torch.logical_or(z_far <= z_near, validity == False)  # noqa: E712
    z_near = torch.where(mask, torch.zeros_like(z_near), z_near)
    z_far = torch.where(mask, torch.zeros_like(z_far), z_far)
    return z_near, z_far


def create_ray_positions(origins: Tensor, directions: Tensor, z_near: Tensor, z_far: Tensor, positions_count: int,
                         perturb: bool, rand: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """utils/lib_3d/ray_helper.py:1229-1282.  ``rand`` replaces ``torch.rand`` of
    line 1275 (uniform [0,1) of shape (..., R, P)) so runs are reproducible."""
    s = torch.linspace(0.0, 1.0, positions_count)
    t = z_near.unsqueeze(-1) * (1.0 - s) + z_far.unsqueeze(-1) * s
    if perturb:
        mid = (t[..., 1:] + t[..., :-1]) / 2
        upper = torch.cat([mid, t[..., -1:]], dim=-1)
        lower = torch.cat([t[..., :1], mid], dim=-1)
        if rand is None:
            rand = torch.rand(t.size())
        t = lower + (upper - lower) * rand
    positions = origins.unsqueeze(-2).unsqueeze(-2) + directions.unsqueeze(-2) * t.unsqueeze(-1)
    return positions, t


def sample_pdf(bin_delimiters: Tensor, weights: Tensor, positions_count: int, perturb: bool, rand: Optional[Tensor] = None) -> Tensor:
    """utils/lib_3d/ray_helper.py:1349-1403: inverse-CDF samples of the piecewise-constant density ``weights`` (..., B-1) over the
    bins delimited by ``bin_delimiters`` (..., B).  ``rand`` replaces the ``torch.rand`` of :1379."""
    weights = weights + 1e-5                                                        # :1365
    pdf = weights / torch.sum(weights, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    lead = list(cdf.shape[:-1])
    if not perturb:
        u = torch.linspace(0.0, 1.0, positions_count).expand(lead + [positions_count]).contiguous()   # :1371-1377
    else:
        u = torch.rand(lead + [positions_count]) if rand is None else rand
    idx = torch.searchsorted(cdf, u, right=True)                                    # :1382
    below = torch.clamp(idx - 1, min=0)
    above = torch.clamp(idx, max=cdf.size(-1) - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_lo, bin_hi = torch.gather(bin_delimiters, -1, below), torch.gather(bin_delimiters, -1, above)
    norm = cdf_hi - cdf_lo
    norm = torch.where(norm < 1e-5, torch.ones_like(norm), norm)                    # :1398
    return bin_lo + (u - cdf_lo) / norm * (bin_hi - bin_lo)


def create_ray_positions_weighted(origins: Tensor, directions: Tensor, positions_count: int, reference_t: Tensor, weights: Tensor,
                                  perturb: bool, rand: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """utils/lib_3d/ray_helper.py:1320-1347: new samples from the coarse weights, merged with the coarse ones and sorted."""
    mid = (reference_t[..., 1:] + reference_t[..., :-1]) / 2
    t_new = sample_pdf(mid, weights[..., 1:-1], positions_count, perturb, rand).detach()
    merged, _ = torch.sort(torch.cat([reference_t, t_new], dim=-1), dim=-1)
    positions = origins.unsqueeze(-2).unsqueeze(-2) + directions.unsqueeze(-2) * merged.unsqueeze(-1)
    return positions, merged


# ----------------------------------------------------------------------------
# Encoders (model/positional_encoder.py, model/annealable_positional_encoder.py)
# ----------------------------------------------------------------------------

def positional_encoding(x: Tensor, octaves: int, append_original: bool = True, weights: Optional[Tensor] = None) -> Tensor:
    """model/positional_encoder.py:41-65 and annealable_positional_encoder.py:46-76.
    Layout: [x, sin(1x), cos(1x), sin(2x), cos(2x), ...], no pi factor."""
    parts: List[Tensor] = [x] if append_original else []
    for k in range(octaves):
        freq = 2.0 ** k
        for fn in (torch.sin, torch.cos):
            e = fn(freq * x)
            if weights is not None:
                e = e * weights[k]
            parts.append(e)
    return torch.cat(parts, dim=-1)


def annealing_weights(current_step: int, octaves: int, num_steps: int) -> Tensor:
    """model/annealable_positional_encoder.py:54-58."""
    alpha = torch.tensor(float(current_step)) * octaves / num_steps
    idx = torch.arange(octaves, dtype=torch.get_default_dtype())
    return (1 - torch.cos(math.pi * torch.clamp(alpha - idx, min=0.0, max=1.0))) / 2


def bounding_box_mask(x: Tensor, bbox: Tensor) -> Tensor:
    """model/nerf_models/ray_bending_style_nerf_model.py:62-85 (inclusive bounds)."""
    return torch.logical_and(x >= bbox[:, 0], x <= bbox[:, 1]).all(dim=-1)


# ----------------------------------------------------------------------------
# Fields (model/nerf_models/*.py, model/layers/adain.py)
# ----------------------------------------------------------------------------

def _linear(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def affine_adain(sd: Dict[str, Tensor], prefix: str, x: Tensor, style: Tensor, training: bool,
                 new_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """model/layers/adain.py:21-36,51-61.  BatchNorm1d(affine=False) then
    per-sample scale/bias from Linear(style).  Train mode: biased batch
    variance for normalisation, unbiased for the running update (recorded in
    ``new_stats`` rather than mutating ``sd``)."""
    enc = _linear(sd, prefix + ".affine_transform", style)
    scale, bias = enc.chunk(2, 1)
    rm = sd[prefix + ".ada_in.normalization.running_mean"]
    rv = sd[prefix + ".ada_in.normalization.running_var"]
    if training and new_stats is not None:
        # a model shared by several object instances is called once per instance: each call updates the statistics the previous one left
        rm = new_stats.get(prefix + ".ada_in.normalization.running_mean", rm)
        rv = new_stats.get(prefix + ".ada_in.normalization.running_var", rv)
    if training:
        if x.size(0) == 1:
            raise ValueError("Expected more than 1 value per channel when training")
        if x.size(0) == 0:
            xn = x
        else:
            mean = x.mean(dim=0)
            var = x.var(dim=0, unbiased=False)
            xn = (x - mean) / torch.sqrt(var + BN_EPS)
            if new_stats is not None:
                n = x.size(0)
                new_stats[prefix + ".ada_in.normalization.running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
                new_stats[prefix + ".ada_in.normalization.running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * n / (n - 1)
    else:
        xn = (x - rm) / torch.sqrt(rv + BN_EPS)
    return xn * scale + bias


def features_head(sd: Dict[str, Tensor], prefix: str, h: Tensor, style: Tensor, training: bool,
                  new_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """adain_style_nerf_model.py:57-71 + adain_sequential.py:14-28:
    Linear(no bias) -> AdaIn -> ReLU -> Linear(no bias) -> AdaIn -> ReLU -> Linear."""
    x = _linear(sd, prefix + ".0", h)
    x = F.relu(affine_adain(sd, prefix + ".1", x, style, training, new_stats))
    x = _linear(sd, prefix + ".3", x)
    x = F.relu(affine_adain(sd, prefix + ".4", x, style, training, new_stats))
    return _linear(sd, prefix + ".6", x)


def _backbone(sd: Dict[str, Tensor], prefix: str, layers: int, skip: int, enc: Tensor, first: Tensor) -> Tensor:
    h = first
    for i in range(layers):
        if i == skip:
            h = torch.cat([h, enc], dim=-1)
        h = F.relu(_linear(sd, f"{prefix}.{i}", h))
    return h


def adain_style_nerf(sd: Dict[str, Tensor], prefix: str, cfg: dict, bbox: Tensor, positions: Tensor, style: Tensor,
                     training: bool, new_stats=None) -> Tuple[Tensor, Tensor]:
    """model/nerf_models/adain_style_nerf_model.py:106-199 on flat (N,3)
    positions, including its own second bounding-box mask (lines 171-184)."""
    n = positions.size(0)
    feats = torch.zeros((n, cfg["output_features"]), dtype=positions.dtype)
    alphas = torch.ones((n, 1), dtype=positions.dtype) * cfg["empty_space_alpha"]
    mask = bounding_box_mask(positions, bbox)
    x = positions[mask] / (bbox[:, 1] - bbox[:, 0])
    enc = positional_encoding(x, cfg["position_encoder"]["octaves"], cfg["position_encoder"]["append_original"])
    h = _backbone(sd, prefix + ".backbone_layers", cfg["backbone_layers_count"], cfg["skip_layer_idx"], enc, enc)
    a = _linear(sd, prefix + ".alpha_head", h)
    f = features_head(sd, prefix + ".features_head", h, style[mask], training, new_stats)
    feats = feats.index_put((mask,), f)
    alphas = alphas.index_put((mask,), a)
    return feats, alphas.squeeze(-1)


def skybox_adain_style_nerf(sd: Dict[str, Tensor], prefix: str, cfg: dict, bbox: Tensor, origins: Tensor,
                            directions: Tensor, style: Tensor, training: bool, new_stats=None) -> Tuple[Tensor, Tensor]:
    """model/nerf_models/skybox_adain_style_nerf_model_v3.py:74-116: input is
    PE(origin/size || unit direction); alpha is forced to 10.0; no mask."""
    size = bbox[:, 1] - bbox[:, 0]
    o = origins / size
    d = directions / directions.pow(2).sum(-1, keepdim=True).sqrt()
    enc = positional_encoding(torch.cat([o, d], dim=-1), cfg["position_encoder"]["octaves"], cfg["position_encoder"]["append_original"])
    h = _backbone(sd, prefix + ".backbone_layers", cfg["backbone_layers_count"], cfg["skip_layer_idx"], enc, enc)
    f = features_head(sd, prefix + ".features_head", h, style, training, new_stats)
    a = torch.ones_like(f[..., 0]) * 10.0
    return f, a


def positional_ray_bender(sd: Dict[str, Tensor], prefix: str, cfg: dict, bbox: Tensor, positions: Tensor,
                          deformation: Tensor, current_step: int) -> Tensor:
    """model/nerf_models/positional_ray_bender_model.py:81-163."""
    size = bbox[:, 1] - bbox[:, 0]
    pe = cfg["position_encoder"]
    w = annealing_weights(current_step, pe["octaves"], pe["num_steps"])
    enc = positional_encoding(positions / size, pe["octaves"], pe["append_original"], w)
    inp = torch.cat([enc, deformation], dim=-1)
    h = _backbone(sd, prefix + ".backbone_layers", cfg["layers_count"], cfg["skip_layer_idx"], inp, inp)
    disp = F.linear(h, sd[prefix + ".output_head.weight"]) * size
    disp = torch.maximum(disp, bbox[:, 0].unsqueeze(0) - positions)       # clamp_output :116-140
    disp = torch.minimum(disp, bbox[:, 1].unsqueeze(0) - positions)
    return disp


def _arch_kind(name: str) -> str:
    return name.rsplit(".", 1)[-1]


def ray_bending_style_nerf(sd: Dict[str, Tensor], cfg: dict, positions: Tensor, origins: Tensor, directions: Tensor,
                           style: Tensor, deformation: Tensor, canonical_pose: bool, training: bool,
                           new_stats=None) -> Tuple[Tensor, Tensor, Tensor]:
    """model/nerf_models/ray_bending_style_nerf_model.py:137-219.
    positions (..., R, P, 3); origins/directions (..., R, 3); style (..., 1|R, S)."""
    bbox = torch.as_tensor(cfg["bounding_box"], dtype=positions.dtype)
    lead = list(positions.shape[:-1])
    P = positions.size(-2)
    flat_pos = positions.reshape(-1, 3)
    flat_org = origins.unsqueeze(-2).expand(lead + [3]).reshape(-1, 3)
    flat_dir = directions.unsqueeze(-2).expand(lead + [3]).reshape(-1, 3)
    flat_style = style.unsqueeze(-2).expand(lead + [style.size(-1)]).reshape(-1, style.size(-1))
    flat_def = deformation.unsqueeze(-2).expand(lead + [deformation.size(-1)]).reshape(-1, deformation.size(-1))
    n = flat_pos.size(0)
    ncfg, bcfg = dict(cfg["nerf_model"]), dict(cfg["ray_bender_model"])
    for c in (ncfg, bcfg):                                    # :39-50
        c["empty_space_alpha"] = cfg["empty_space_alpha"]
    out_f = torch.zeros((n, ncfg["output_features"]), dtype=positions.dtype)
    out_a = torch.ones((n,), dtype=positions.dtype) * cfg["empty_space_alpha"]
    out_d = torch.zeros((n, 3), dtype=positions.dtype)
    mask = bounding_box_mask(flat_pos, bbox)
    pos, sty, dfm = flat_pos[mask], flat_style[mask], flat_def[mask]

    bender = _arch_kind(bcfg["architecture"])
    if bender == "zeroed_ray_bender_model":
        disp = pos * 0.0                                      # zeroed_ray_bender_model.py:28-37
    elif bender == "positional_ray_bender_model":
        disp = positional_ray_bender(sd, "ray_bender", bcfg, bbox, pos, dfm, int(sd["ray_bender.positional_encoder.current_step"]))
    else:
        raise Exception(f"oracle: unsupported ray bender {bender}")
    if canonical_pose:
        disp = disp * 0.0
    bent = pos + disp

    nerf = _arch_kind(ncfg["architecture"])
    if nerf == "adain_style_nerf_model":
        f, a = adain_style_nerf(sd, "nerf_model", ncfg, bbox, bent, sty, training, new_stats)
    elif nerf == "skybox_adain_style_nerf_model_v3":
        f, a = skybox_adain_style_nerf(sd, "nerf_model", ncfg, bbox, flat_org[mask], flat_dir[mask], sty, training, new_stats)
    else:
        raise Exception(f"oracle: unsupported nerf model {nerf}")
    out_f = out_f.index_put((mask,), f)
    out_a = out_a.index_put((mask,), a)
    out_d = out_d.index_put((mask,), disp)
    F_out = ncfg["output_features"]
    return out_f.reshape(lead + [F_out]), out_a.reshape(lead), out_d.reshape(lead + [3])


# ----------------------------------------------------------------------------
# Compositing (model/object_composer.py)
# ----------------------------------------------------------------------------

def position_distances(t: Tensor, directions: Tensor) -> Tensor:
    """model/object_composer.py:153-178: last interval 1e10, scaled by |d|."""
    first = t[..., 1:] - t[..., :-1]
    last = torch.ones_like(t[..., :1]) * 1e10
    return torch.cat([first, last], dim=-1) * torch.linalg.norm(directions[..., None, :], dim=-1)


def compute_alphas(raw: Tensor, dist: Tensor, noise: Optional[Tensor]) -> Tensor:
    """model/object_composer.py:180-197; ``noise`` replaces torch.randn."""
    if noise is not None:
        raw = raw + noise
    return 1.0 - torch.exp(-F.relu(raw) * dist)


def compute_weights(alphas: Tensor) -> Tensor:
    """model/object_composer.py:199-214: exclusive cumprod of (1 - a + 1e-10)."""
    shift = 1.0 - alphas + 1e-10
    shift = torch.cat([torch.ones_like(shift[..., :1]), shift[..., :-1]], dim=-1)
    return alphas * torch.cumprod(shift, dim=-1)


def integrate(features: Tensor, raw_alphas: Tensor, directions: Tensor, t: Tensor, displacements: Tensor,
              divergences: Tensor, noise: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """model/object_composer.py:724-784."""
    dist = position_distances(t, directions)
    alphas = compute_alphas(raw_alphas, dist, noise)
    weights = compute_weights(alphas)
    integrated = torch.sum(weights.unsqueeze(-1) * features, dim=-2)
    depth = torch.sum(weights * t, dim=-1)
    opacity = torch.sum(weights, dim=-1)
    disparity = 1.0 / torch.clamp(depth / opacity, min=1e-10)
    int_div = torch.mean(alphas.detach() * torch.abs(divergences), dim=-1)
    int_disp = torch.mean(weights.detach() * torch.norm(displacements, dim=-1), dim=-1)
    return {
        "integrated_features": integrated, "opacity": opacity, "weights": weights, "depth": depth,
        "disparity": disparity, "integrated_displacements_magnitude": int_disp, "integrated_divergence": int_div,
    }


def fix_object_overlap_mask(original_static_t: Tensor, dynamic_t: Tensor) -> Tensor:
    """model/object_composer.py:295-360, vectorised: static samples with index in
    [searchsorted(t_static, t_dyn[0]), searchsorted(t_static, t_dyn[P_static-1]))
    are masked.  NB the reference indexes the dynamic t at ``positions_count - 1``
    of the *static* object (line 322)."""
    P = original_static_t.size(-1)
    bounds = dynamic_t[..., (0, P - 1)]
    iv = torch.searchsorted(original_static_t.contiguous(), bounds.contiguous())
    idx = torch.arange(P).expand_as(original_static_t)
    return torch.logical_and(idx >= iv[..., :1], idx < iv[..., 1:])


def compose(fix_overlaps: bool, static_count: int, ray_origins_exp: Tensor, feats: List[Tensor], raws: List[Tensor],
            ts: List[Tensor], poss: List[Tensor], disps: List[Tensor], divs: List[Tensor]):
    """model/object_composer.py:399-447 (+220-397 when fix_object_overlaps)."""
    if fix_overlaps:
        raws, ts, poss, disps, divs = list(raws), list(ts), list(poss), list(disps), list(divs)
        orig_ts = list(ts)
        for s in range(static_count):
            for d in range(static_count, len(raws)):
                m = fix_object_overlap_mask(orig_ts[s], orig_ts[d])
                raws[s] = torch.where(m, raws[s] * 0.0 - 10.0, raws[s])
                ts[s] = torch.where(m, ts[s] * 0.0, ts[s])
                poss[s] = torch.where(m.unsqueeze(-1), ray_origins_exp.unsqueeze(-2).expand_as(poss[s]), poss[s])
                disps[s] = torch.where(m.unsqueeze(-1), disps[s] * 0.0, disps[s])
                divs[s] = torch.where(m, divs[s] * 0.0, divs[s])
    f = torch.cat(feats, dim=-2)
    a = torch.cat(raws, dim=-1)
    t = torch.cat(ts, dim=-1)
    p = torch.cat(poss, dim=-2)
    d = torch.cat(disps, dim=-2)
    v = torch.cat(divs, dim=-1)
    # torch.sort at :435 is unstable; stable=True fixes a deterministic tie order
    # (object index, then sample index) that the CUDA path reproduces.
    t, idx = torch.sort(t, dim=-1, stable=True)
    a = torch.gather(a, -1, idx)
    v = torch.gather(v, -1, idx)
    f = torch.gather(f, -2, idx.unsqueeze(-1).expand_as(f))
    p = torch.gather(p, -2, idx.unsqueeze(-1).expand_as(p))
    d = torch.gather(d, -2, idx.unsqueeze(-1).expand_as(d))
    return f, a, t, p, d, v


def object_ids(config: dict) -> Tuple[List[int], int]:
    """model/utils/object_ids_helper.py:4-45: object instance -> model index,
    static models first; returns (model_idx per object, static object count)."""
    m = config["model"]
    model_of: List[int] = []
    static = 0
    for mi in range(len(m["object_models"])):
        cnt = m["object_parameters_encoder"][mi]["objects_count"]
        for _ in range(cnt):
            model_of.append(mi)
            if mi < m["static_object_models"]:
                static += 1
    return model_of, static


def composer_forward(config: dict, state: Dict[str, Tensor], ray_origins: Tensor, ray_directions: Tensor,
                     focal_normals: Tensor, transformation_matrix_w2o: Tensor, style: Tensor, deformation: Tensor,
                     object_in_scene: Tensor, perturb: bool, canonical_pose: bool = False, training: bool = False,
                     rand: Optional[List[Tensor]] = None, noise: Optional[Dict[str, Tensor]] = None,
                     new_stats: Optional[Dict[str, Tensor]] = None, divergence_noise: Optional[List[Tensor]] = None) -> Dict:
    """model/object_composer.py:786-892 (+ forward_object :486-580); with ``use_fine`` object models also the fine pass (:561-578:
    ``object_models_fine.{m}.<param>`` on the coarse samples merged with inverse-CDF samples of the coarse weights).

    ``state`` holds ``object_models_coarse.{m}.<param>`` tensors.  ``rand[k]``
    (uniform, shape (..., R, P_k)) and ``noise["object_k"|"global"]`` (normal)
    stand in for the reference's RNG calls when ``perturb`` is set.  The
    Hutchinson divergence (:582-601) is random by construction: it is evaluated
    when its probe vectors ``divergence_noise[k]`` (..., R, P_k, 3) are given
    (training only, like the reference), else returned as zeros."""
    m = config["model"]
    model_of, static_count = object_ids(config)
    objects_count = len(model_of)
    if transformation_matrix_w2o.size(-1) != objects_count:
        raise Exception(f"Transformation matrix must specifies transformations for({transformation_matrix_w2o.size(-1)}) objects instead of ({objects_count})")
    R = ray_directions.size(-2)
    per_obj = []
    per_obj_fine = []
    use_fine = m["object_models"][model_of[0]].get("use_fine", True)
    for k in range(objects_count):
        mi = model_of[k]
        cfg = m["object_models"][mi]
        prefix = f"object_models_coarse.{mi}."
        sd = {key[len(prefix):]: val for key, val in state.items() if key.startswith(prefix)}
        bbox = torch.as_tensor(cfg["bounding_box"], dtype=ray_directions.dtype)
        w2o = transformation_matrix_w2o[..., k]
        ois = object_in_scene[..., k]
        o, d, _ = transform_rays(ray_origins, ray_directions, focal_normals, w2o)
        zn, zf = raywise_object_z_bounds(o, d, bbox, ois)
        zn = torch.clamp(zn, min=cfg["z_near_min"], max=cfg["z_far_max"])
        zf = torch.clamp(zf, min=cfg["z_near_min"], max=cfg["z_far_max"])
        P = cfg["positions_count_coarse"]
        pos, t = create_ray_positions(o, d, zn, zf, P, perturb, None if rand is None else rand[k])
        o_exp = o.unsqueeze(-2).expand(list(d.shape))
        want_div = training and divergence_noise is not None and divergence_noise[k] is not None
        if want_div:
            pos = pos.detach().requires_grad_(True)
        with torch.enable_grad() if want_div else contextlib.nullcontext():
            f, a, disp = ray_bending_style_nerf(sd, cfg, pos, o_exp, d, style[..., k].unsqueeze(-2), deformation[..., k].unsqueeze(-2),
                                                canonical_pose, training, None if new_stats is None else _Prefixed(new_stats, prefix))
        div = torch.zeros_like(a)
        if want_div:
            # compute_approximate_divergence :582-601: e^T (d displacement / d position) e by one vector-Jacobian product
            e = divergence_noise[k]
            e_dydx = torch.autograd.grad(disp, pos, e, allow_unused=True)[0] if disp.requires_grad else None
            div = (e_dydx * e).sum(dim=-1) if e_dydx is not None else div
            f, a, disp, pos, div = f.detach(), a.detach(), disp.detach(), pos.detach(), div.detach()
        absent = torch.logical_not(ois)
        a = torch.where(absent.reshape(list(absent.shape) + [1] * (a.dim() - absent.dim())), torch.full_like(a, cfg["empty_space_alpha"]), a)
        if m["apply_activation"]:
            f = torch.sigmoid(f)
        per_obj.append((f, a, t, pos, disp, div))
        if use_fine:
            # :552-578 -- coarse weights of THIS object (no raw-alpha noise here: the fine goldens are unperturbed), resampling, fine model
            w_c = compute_weights(compute_alphas(a, position_distances(t, d), None))
            fprefix = f"object_models_fine.{mi}."
            fsd = {key[len(fprefix):]: val for key, val in state.items() if key.startswith(fprefix)}
            fpos, ft = create_ray_positions_weighted(o, d, cfg["positions_count_fine"], t, w_c, perturb)
            ff, fa, fdisp = ray_bending_style_nerf(fsd, cfg, fpos, o_exp, d, style[..., k].unsqueeze(-2), deformation[..., k].unsqueeze(-2),
                                                   canonical_pose, training, None if new_stats is None else _Prefixed(new_stats, fprefix))
            fa = torch.where(absent.reshape(list(absent.shape) + [1] * (fa.dim() - absent.dim())), torch.full_like(fa, cfg["empty_space_alpha"]), fa)
            if m["apply_activation"]:
                ff = torch.sigmoid(ff)
            per_obj_fine.append((ff, fa, ft, fpos, fdisp, torch.zeros_like(fa)))

    results: Dict = {}
    exp_origins = ray_origins.unsqueeze(-2).expand(list(ray_directions.shape))
    for model_type, objs in (("coarse", per_obj), ("fine", per_obj_fine)):
        if not objs:
            continue
        results[model_type] = {}
        for k, (f, a, t, pos, disp, div) in enumerate(objs):
            nz = None if (noise is None or not perturb or model_type == "fine") else noise[f"object_{k}"]
            r = integrate(f, a, ray_directions, t, disp, div, nz)
            r["extra_outputs"] = {}
            results[model_type][f"object_{k}"] = r
        cf, ca, ct, cp, cd, cv = compose(m.get("fix_object_overlaps", True), static_count, exp_origins,
                                         [x[0] for x in objs], [x[1] for x in objs], [x[2] for x in objs],
                                         [x[3] for x in objs], [x[4] for x in objs], [x[5] for x in objs])
        nz = None if (noise is None or not perturb or model_type == "fine") else noise["global"]
        results[model_type]["global"] = integrate(cf, ca, ray_directions, ct, cd, cv, nz)
    results["pytorch_hook"] = torch.zeros((1,) * 9)
    return results


def forward_expected_positions(config: dict, state: Dict[str, Tensor], ray_origins: Tensor, ray_directions: Tensor, focal_normals: Tensor,
                               transformation_matrix_w2o: Tensor, style: Tensor, deformation: Tensor, object_in_scene: Tensor,
                               object_id: int, perturb: bool, canonical_pose: bool = False, training: bool = False,
                               rand: Optional[Tensor] = None, noise: Optional[Tensor] = None) -> Dict:
    """model/object_composer.py:624-722 (coarse pass) + compute_expected_positions :603-622, for ONE object instance:
    ``transformation_matrix_w2o`` (..., 4, 4), ``style`` (..., S), ``deformation`` (..., D), ``object_in_scene`` (...).
    Returns {"coarse": (expected_positions (..., R, 3) in object space, opacity (..., R))}."""
    model_of, _ = object_ids(config)
    mi = model_of[object_id]
    cfg = config["model"]["object_models"][mi]
    prefix = f"object_models_coarse.{mi}."
    sd = {key[len(prefix):]: val for key, val in state.items() if key.startswith(prefix)}
    bbox = torch.as_tensor(cfg["bounding_box"], dtype=torch.float32)
    o, d, _ = transform_rays(ray_origins, ray_directions, focal_normals, transformation_matrix_w2o)
    zn, zf = raywise_object_z_bounds(o, d, bbox, object_in_scene)
    zn = torch.clamp(zn, min=cfg["z_near_min"], max=cfg["z_far_max"])
    zf = torch.clamp(zf, min=cfg["z_near_min"], max=cfg["z_far_max"])
    pos, t = create_ray_positions(o, d, zn, zf, cfg["positions_count_coarse"], perturb, rand)
    o_exp = o.unsqueeze(-2).expand(list(d.shape))
    _, a, disp = ray_bending_style_nerf(sd, cfg, pos, o_exp, d, style.unsqueeze(-2), deformation.unsqueeze(-2), canonical_pose, training)
    absent = torch.logical_not(object_in_scene)
    a = torch.where(absent.reshape(list(absent.shape) + [1] * (a.dim() - absent.dim())), torch.full_like(a, cfg["empty_space_alpha"]), a)
    # :687 -- the reference has re-bound ``ray_directions`` to the OBJECT-space directions by now (:656), so the sample spacing is
    # scaled by |R d| here (forward() scales by the world-space |d|, :880): the same value for a rigid pose, but its gradient w.r.t.
    # the rotation block has the extra component R d d^T / |d|
    weights = compute_weights(compute_alphas(a, position_distances(t, d), noise if perturb else None))
    w = weights.detach().unsqueeze(-1)
    expected = ((pos + disp) * w).sum(dim=-2) / (w.sum(dim=-2) + 1e-8)          # :603-622
    return {"coarse": (expected, weights.sum(dim=-1))}


class _Prefixed(dict):
    """dict view that writes ``prefix + key`` into a parent dict."""

    def __init__(self, parent: Dict[str, Tensor], prefix: str):
        super().__init__()
        self._parent, self._prefix = parent, prefix

    def __setitem__(self, key, value):
        self._parent[self._prefix + key] = value

    def get(self, key, default=None):
        return self._parent.get(self._prefix + key, default)


# ----------------------------------------------------------------------------
# Caller-side helpers on the path
# ----------------------------------------------------------------------------

def batchify(tensor: Tensor, dim: int, batch_size: int) -> List[Tensor]:
    """utils/tensor_batchifier.py:9-45."""
    if dim < 0:
        dim += tensor.dim()
    return list(torch.split(tensor, batch_size, dim=dim))


def merge_dictionaries(dicts: List[Dict], dimension: int) -> Dict:
    """model/environment_model.py:523-545 (drops ``pytorch_hook``)."""
    out: Dict = {}
    for key in dicts[0]:
        if key == "pytorch_hook":
            continue
        if torch.is_tensor(dicts[0][key]):
            out[key] = torch.cat([d[key] for d in dicts], dim=dimension)
        else:
            out[key] = merge_dictionaries([d[key] for d in dicts], dimension)
    return out


def batchified_composer_call(config, state, ray_origins, ray_directions, focal_normals, w2o, style, deformation,
                             object_in_scene, perturb, samples_per_image_batching: int = 0, **kw) -> Dict:
    """model/environment_model.py:474-521."""
    dim = ray_directions.dim() - 2
    bs = samples_per_image_batching or ray_directions.size(dim)
    chunks = batchify(ray_directions, -2, bs)
    res = [composer_forward(config, state, ray_origins, c, focal_normals, w2o, style, deformation, object_in_scene, perturb, **kw) for c in chunks]
    return merge_dictionaries(res, dim)


def fold_strided_grid_samples(samples: Tensor, strides, original_size: Tuple[int, int], dim: int) -> List[Tensor]:
    """utils/lib_3d/ray_helper.py:484-531."""
    if not isinstance(strides, (list, tuple)):
        strides = [strides]
    H, W = original_size
    out, start = [], 0
    for s in strides:
        gh, gw = H // s, W // s
        sl = [slice(None)] * samples.dim()
        sl[dim] = slice(start, start + gh * gw)
        cur = samples[tuple(sl)]
        shape = list(cur.shape)
        shape[dim:dim + 1] = [gh, gw]
        out.append(cur.reshape(shape))
        start += gh * gw
    return out


def decoder_feature_grids(integrated_features: Tensor, strides: Sequence[int], original_size: Tuple[int, int],
                          features_per_stride: Sequence[int]) -> List[Tensor]:
    """environment_model_backpropagated_autoencoder.py:129-168 followed by
    environment_model_multiresolution_backpropagated_autoencoder.py:59-99:
    (..., R, F) -> per-stride CHW grids keeping a disjoint channel range each
    (e.g. 0:64 at stride 4 and 64:192 at stride 8)."""
    grids = fold_strided_grid_samples(integrated_features, list(strides), original_size, dim=-2)
    out, c0 = [], 0
    for g, nf in zip(grids, features_per_stride):
        g = g[..., c0:c0 + nf]
        out.append(g.movedim(-1, -3).contiguous())
        c0 += nf
    return out
