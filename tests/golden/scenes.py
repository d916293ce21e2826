"""Seeded synthetic scenes shared by the golden generator, the tests and bench.py.

Everything here is derived from ``numpy.random.default_rng(seed)`` so the same
weights and inputs are rebuilt bit-identically on any box without shipping
megabytes of parameters; only the *reference outputs* are stored (``*.npz``
next to this file, written by ``make_golden.py`` which imports the upstream
code).  Config dictionaries use the reference's YAML keys verbatim
(configs/tennis/193_*.yaml:96-330, configs/minecraft/013_*.yaml:56-290).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

REF_PREFIX = "model.nerf_models."


# ----------------------------------------------------------------------------
# object-model configs
# ----------------------------------------------------------------------------

def nerf_cfg(width=256, layers=8, skip=4, octaves=10, features=192, kind="adain_style_nerf_model") -> dict:
    return {
        "architecture": REF_PREFIX + kind,
        "layers_width": width, "backbone_layers_count": layers, "output_features": features, "skip_layer_idx": skip,
        "position_encoder": {"octaves": octaves, "append_original": True},
    }


def bender_cfg(kind="zeroed", width=128, layers=6, skip=3, octaves=6, num_steps=60000) -> dict:
    if kind == "zeroed":
        return {"architecture": REF_PREFIX + "zeroed_ray_bender_model"}
    return {
        "architecture": REF_PREFIX + "positional_ray_bender_model",
        "layers_width": width, "layers_count": layers, "skip_layer_idx": skip,
        "position_encoder": {"octaves": octaves, "append_original": True, "num_steps": num_steps},
    }


def object_cfg(bbox, P, z_near_min, z_far_max, style, deformation, nerf, bender, empty_space_alpha=-3.5) -> dict:
    return {
        "architecture": REF_PREFIX + "ray_bending_style_nerf_model",
        "bounding_box": [list(map(float, b)) for b in bbox],
        "positions_count_coarse": P, "positions_count_fine": P, "use_fine": False,
        "empty_space_alpha": empty_space_alpha, "z_near_min": z_near_min, "z_far_max": z_far_max,
        "deformation_features": deformation, "style_features": style,
        "nerf_model": nerf, "ray_bender_model": bender,
    }


def scene_config(object_models: List[dict], static_models: int, objects_per_model: List[int], fix_overlaps: bool) -> dict:
    return {"model": {
        "apply_activation": False, "fix_object_overlaps": fix_overlaps, "static_object_models": static_models,
        "object_parameters_encoder": [{"objects_count": c} for c in objects_per_model],
        "object_encoders": [{} for _ in objects_per_model],
        "object_models": object_models,
    }}


# ----------------------------------------------------------------------------
# seeded parameters (reference state_dict names)
# ----------------------------------------------------------------------------

def _uniform(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def _linear(rng, sd, name, fan_in, fan_out, bias=True, bound=None):
    b = 1.0 / math.sqrt(fan_in) if bound is None else bound
    sd[name + ".weight"] = _uniform(rng, (fan_out, fan_in), b)
    if bias:
        sd[name + ".bias"] = _uniform(rng, (fan_out,), 1.0 / math.sqrt(fan_in))


def _adain(rng, sd, name, channels, style):
    _linear(rng, sd, name + ".affine_transform", style, 2 * channels)
    sd[name + ".affine_transform.bias"][:channels] += 1.0          # scale biased to 1 (adain.py:17-19)
    sd[name + ".ada_in.normalization.running_mean"] = torch.from_numpy(rng.normal(0.0, 0.1, channels).astype(np.float32))
    sd[name + ".ada_in.normalization.running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, channels).astype(np.float32))
    sd[name + ".ada_in.normalization.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def object_state(rng, cfg: dict, current_step: int = 60000) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    n = cfg["nerf_model"]
    W, L, Fo, S = n["layers_width"], n["backbone_layers_count"], n["output_features"], cfg["style_features"]
    in_dims = 6 if n["architecture"].endswith("skybox_adain_style_nerf_model_v3") else 3
    E = in_dims * (1 + 2 * n["position_encoder"]["octaves"])
    cur = E
    for i in range(L):
        if i == n["skip_layer_idx"]:
            cur += E
        # He-like bound keeps activations O(1) through 8 layers so parity is not vacuous
        _linear(rng, sd, f"nerf_model.backbone_layers.{i}", cur, W, bound=math.sqrt(6.0 / cur))
        sd[f"nerf_model.backbone_layers.{i}.bias"] = _uniform(rng, (W,), 0.1)
        cur = W
    if not n["architecture"].endswith("skybox_adain_style_nerf_model_v3"):
        _linear(rng, sd, "nerf_model.alpha_head", W, 1, bound=math.sqrt(6.0 / W))
        sd["nerf_model.alpha_head.bias"] = torch.tensor([0.5])        # SURVEY 8c caveat: make opacity non-trivial
    _linear(rng, sd, "nerf_model.features_head.0", W, W, bias=False, bound=math.sqrt(6.0 / W))
    _adain(rng, sd, "nerf_model.features_head.1", W, S)
    _linear(rng, sd, "nerf_model.features_head.3", W, W // 2, bias=False, bound=math.sqrt(6.0 / W))
    _adain(rng, sd, "nerf_model.features_head.4", W // 2, S)
    _linear(rng, sd, "nerf_model.features_head.6", W // 2, Fo, bound=math.sqrt(6.0 / (W // 2)))
    b = cfg["ray_bender_model"]
    if b["architecture"].endswith("positional_ray_bender_model"):
        Wb, Lb = b["layers_width"], b["layers_count"]
        Eb = 3 * (1 + 2 * b["position_encoder"]["octaves"]) + cfg["deformation_features"]
        cur = Eb
        for i in range(Lb):
            if i == b["skip_layer_idx"]:
                cur += Eb
            _linear(rng, sd, f"ray_bender.backbone_layers.{i}", cur, Wb, bound=math.sqrt(6.0 / cur))
            sd[f"ray_bender.backbone_layers.{i}.bias"] = _uniform(rng, (Wb,), 0.05)
            cur = Wb
        # visible displacements (the reference inits these near zero, which would hide bender bugs)
        sd["ray_bender.output_head.weight"] = _uniform(rng, (3, Wb), 0.02)
        sd["ray_bender.positional_encoder.current_step"] = torch.tensor(current_step, dtype=torch.int32)
    return sd


def scene_state(seed: int, config: dict, current_step: int = 60000) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    state: Dict[str, torch.Tensor] = {}
    for mi, cfg in enumerate(config["model"]["object_models"]):
        for k, v in object_state(rng, cfg, current_step).items():
            state[f"object_models_coarse.{mi}.{k}"] = v
    for mi, cfg in enumerate(config["model"]["object_models"]):        # fine models (object_composer.py:27, 44-53) have their own weights
        if cfg.get("use_fine", True):
            for k, v in object_state(rng, cfg, current_step).items():
                state[f"object_models_fine.{mi}.{k}"] = v
    return state


# ----------------------------------------------------------------------------
# cameras and inputs
# ----------------------------------------------------------------------------

def rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def homogeneous(rotation3x3, translation) -> np.ndarray:
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = rotation3x3
    m[:3, 3] = translation
    return m


def camera_rays(lead: Tuple[int, ...], height: int, width: int, focal: float, c2w: np.ndarray, stride: int = 1):
    """Pinhole rays (utils/lib_3d/ray_helper.py:15-52 convention), strided-grid
    selection (:533-582), moved to world space with ``c2w`` (transform_rays :1203)."""
    off = stride // 2
    rows = np.arange(height // stride) * stride + off
    cols = np.arange(width // stride) * stride + off
    rr, cc = np.meshgrid(rows, cols, indexing="ij")
    d = np.stack([(cc - width / 2) / focal, -(rr - height / 2) / focal, -np.ones_like(rr, dtype=np.float64)], -1).reshape(-1, 3)
    d = d @ c2w[:3, :3].T
    o = c2w[:3, 3]
    n = c2w[:3, :3] @ np.array([0.0, 0.0, -1.0])
    R = d.shape[0]
    dirs = torch.from_numpy(np.broadcast_to(d, lead + (R, 3)).astype(np.float32).copy())
    orig = torch.from_numpy(np.broadcast_to(o, lead + (3,)).astype(np.float32).copy())
    norm = torch.from_numpy(np.broadcast_to(n, lead + (3,)).astype(np.float32).copy())
    return orig, dirs, norm


def _codes(rng, lead, size, objects):
    return torch.from_numpy(rng.normal(0.0, 1.0, lead + (size, objects)).astype(np.float32))


def build_inputs(seed: int, config: dict, lead, orig, dirs, norm, w2o_list: List[np.ndarray], absent=()) -> dict:
    """w2o_list: one (4,4) world->object matrix per object instance."""
    rng = np.random.default_rng(seed + 1000)
    objs = len(w2o_list)
    first = config["model"]["object_models"][0]
    S, D = first["style_features"], first["deformation_features"]
    w2o = np.stack(w2o_list, axis=-1).astype(np.float32)                       # (4,4,objs)
    w2o = torch.from_numpy(np.broadcast_to(w2o, lead + w2o.shape).copy())
    ois = torch.ones(lead + (objs,), dtype=torch.bool)
    for idx in absent:
        ois[idx] = False
    return {
        "ray_origins": orig, "ray_directions": dirs, "focal_normals": norm, "transformation_matrix_w2o": w2o,
        "style": _codes(rng, lead, S, objs), "deformation": _codes(rng, lead, D, objs), "object_in_scene": ois,
    }


def perturbation_tensors(seed: int, config: dict, inputs: dict):
    """Uniform ``rand[k]`` (stratified jitter, ray_helper.py:1275) and normal
    ``noise`` (raw-alpha noise, object_composer.py:194) in the call order of the
    reference: rand per object in forward_object, then randn per object
    integrate, then the global one."""
    rng = np.random.default_rng(seed + 2000)
    lead_r = tuple(inputs["ray_directions"].shape[:-1])
    model_of = []
    for mi, c in enumerate(config["model"]["object_parameters_encoder"]):
        model_of += [mi] * c["objects_count"]
    Ps = [config["model"]["object_models"][mi]["positions_count_coarse"] for mi in model_of]
    rand = [torch.from_numpy(rng.uniform(0.0, 1.0, lead_r + (p,)).astype(np.float32)) for p in Ps]
    noise = {f"object_{k}": torch.from_numpy(rng.normal(0, 1, lead_r + (p,)).astype(np.float32)) for k, p in enumerate(Ps)}
    noise["global"] = torch.from_numpy(rng.normal(0, 1, lead_r + (sum(Ps),)).astype(np.float32))
    return rand, noise


def divergence_noise(seed: int, config: dict, inputs: dict):
    """Normal probe vectors ``e`` of the Hutchinson divergence (object_composer.py:597), one (..., R, P_k, 3) tensor per object instance."""
    rng = np.random.default_rng(seed + 3000)
    lead_r = tuple(inputs["ray_directions"].shape[:-1])
    model_of = []
    for mi, c in enumerate(config["model"]["object_parameters_encoder"]):
        model_of += [mi] * c["objects_count"]
    Ps = [config["model"]["object_models"][mi]["positions_count_coarse"] for mi in model_of]
    return [torch.from_numpy(rng.normal(0, 1, lead_r + (p, 3)).astype(np.float32)) for p in Ps]


# ----------------------------------------------------------------------------
# the scenes (BASELINE.json configs)
# ----------------------------------------------------------------------------

def scene_cfg1(seed=11):
    """BASELINE configs[0]: single 4x64 field, 32x32 rays, 16 samples/ray."""
    obj = object_cfg([[-2, 2], [-2, 2], [-6, -2]], 16, 0.05, 30.0, 16, 8,
                     nerf_cfg(64, 4, 2, 4, 3), bender_cfg("zeroed"))
    config = scene_config([obj], 1, [1], False)
    lead = (1, 1, 1)
    orig, dirs, norm = camera_rays(lead, 32, 32, 40.0, np.eye(4))
    inputs = build_inputs(seed, config, lead, orig, dirs, norm, [np.eye(4)])
    return config, scene_state(seed, config), inputs


def scene_static(seed=12, height=16, width=16, P=128, lead=(1, 1, 1), style=32):
    """BASELINE configs[1] shape (Minecraft static field: W=256, L=8, skip 4,
    10 octaves, F=192; camera inside a large box so every sample is in-box)."""
    obj = object_cfg([[-10, 10], [-10, 10], [-10, 10]], P, 0.05, 8.0, style, 32, nerf_cfg(), bender_cfg("zeroed"))
    config = scene_config([obj], 1, [1], True)
    c2w = homogeneous(rot_x(0.2) @ rot_z(0.3), [0.3, -0.2, 0.5])
    orig, dirs, norm = camera_rays(lead, height, width, 0.9 * width, c2w)
    inputs = build_inputs(seed, config, lead, orig, dirs, norm, [np.eye(4)])
    return config, scene_state(seed, config), inputs


def tennis_camera():
    """Broadcast camera of SURVEY 8d cfg3: tilt 1.25 rad, at (0,-32,9)."""
    return homogeneous(rot_x(1.25), [0.0, -32.0, 9.0])


def scene_tennis(seed=13, height=144, width=256, stride=8, lead=(1, 2, 1), step=60000, dense=False):
    """BASELINE configs[2] shape: court (static, P=4) + 2 players (positional
    ray bender, P=32), S=64, D=32.  ``dense`` zooms the camera onto a player so
    most rays traverse a player box."""
    court = object_cfg([[-30, 30], [-40, 20.585], [-0.5, 0.0]], 4, 5.0, 70.0, 64, 32, nerf_cfg(), bender_cfg("zeroed"))
    player = lambda: object_cfg([[-0.75, 0.75], [-0.5, 0.5], [0.0, 2.15]], 32, 5.0, 70.0, 64, 32, nerf_cfg(), bender_cfg("positional"))
    config = scene_config([court, player(), player()], 1, [1, 1, 1], False)
    focal = 1700.0 * 0.51417 * 0.5 * (width / 256.0)
    if dense:
        focal *= 12.0
    c2w = tennis_camera()
    orig, dirs, norm = camera_rays(lead, height, width, focal, c2w, stride)
    p1 = np.linalg.inv(homogeneous(rot_z(0.3), [2.0, -11.0, 0.01]))
    p2 = np.linalg.inv(homogeneous(rot_z(-0.2), [-2.0, 11.0, 0.01]))
    if dense:
        p1 = np.linalg.inv(homogeneous(rot_z(0.3), [0.0, -6.0, 0.01]))
    inputs = build_inputs(seed, config, lead, orig, dirs, norm, [np.eye(4), p1, p2])
    return config, scene_state(seed, config, step), inputs


def scene_minecraft(seed=14, height=64, width=64, stride=4, lead=(1, 1, 1), absent=()):
    """Minecraft-shaped scene: ground (P=16) + skybox (P=1, t in [90,91]) static,
    one player model shared by 2 instances (P=32), fix_object_overlaps on."""
    ground = object_cfg([[-10, 10], [-0.6, 2.0], [-10, 10]], 16, 0.05, 30.0, 32, 32, nerf_cfg(), bender_cfg("zeroed"))
    sky = object_cfg([[-200, 200], [-200, 200], [-200, 200]], 1, 90.0, 91.0, 32, 32,
                     nerf_cfg(kind="skybox_adain_style_nerf_model_v3"), bender_cfg("zeroed"))
    player = object_cfg([[-0.6, 0.6], [0.0, 2.1], [-1.2, 1.2]], 32, 0.05, 30.0, 32, 32, nerf_cfg(), bender_cfg("positional"))
    config = scene_config([ground, sky, player], 2, [1, 1, 2], True)
    c2w = homogeneous(rot_x(-0.25), [0.0, 1.6, 6.0])
    orig, dirs, norm = camera_rays(lead, height, width, 0.8 * width, c2w, stride)
    pa = np.linalg.inv(homogeneous(np.eye(3), [-0.8, 0.0, 1.0]))
    pb = np.linalg.inv(homogeneous(np.eye(3), [1.0, 0.0, -0.5]))
    inputs = build_inputs(seed, config, lead, orig, dirs, norm, [np.eye(4), np.eye(4), pa, pb], absent=absent)
    return config, scene_state(seed, config), inputs


def scene_toy_world(seed=18, height=64, width=64, stride=4, lead=(1, 2, 1)):
    """Every component of the path with small, well-conditioned networks (4 octaves: the gradients are not dominated by fp32
    cancellation, so the backward can be pinned tightly): ground (P=8) + skybox (P=1) static, one player model with a positional
    ray bender shared by 2 instances (P=12), fix_object_overlaps on, two images with different codes."""
    toy = lambda kind="adain_style_nerf_model": nerf_cfg(64, 4, 2, 4, 8, kind)
    ground = object_cfg([[-10, 10], [-0.6, 2.0], [-10, 10]], 8, 0.05, 30.0, 16, 8, toy(), bender_cfg("zeroed"))
    sky = object_cfg([[-200, 200], [-200, 200], [-200, 200]], 1, 90.0, 91.0, 16, 8, toy("skybox_adain_style_nerf_model_v3"), bender_cfg("zeroed"))
    player = object_cfg([[-0.6, 0.6], [0.0, 2.1], [-1.2, 1.2]], 12, 0.05, 30.0, 16, 8, toy(),
                        bender_cfg("positional", width=32, layers=3, skip=1, octaves=3))
    config = scene_config([ground, sky, player], 2, [1, 1, 2], True)
    c2w = homogeneous(rot_x(-0.25), [0.0, 1.6, 6.0])
    orig, dirs, norm = camera_rays(lead, height, width, 0.8 * width, c2w, stride)
    pa = np.linalg.inv(homogeneous(rot_z(0.1), [-0.8, 0.0, 1.0]))
    pb = np.linalg.inv(homogeneous(np.eye(3), [1.0, 0.0, -0.5]))
    inputs = build_inputs(seed, config, lead, orig, dirs, norm, [np.eye(4), np.eye(4), pa, pb])
    state = scene_state(seed, config)
    for k in list(state):                       # larger displacements: part of the samples hit the clamp at the box faces
        if k.endswith("ray_bender.output_head.weight"):
            state[k] = state[k] * 8.0
    return config, state, inputs


def with_fine(scene, fine_counts: List[int]):
    """The same scene with a fine model per object model (use_fine, positions_count_fine = fine_counts[model]); weights re-seeded."""
    config, _, inputs = scene
    for cfg, pf in zip(config["model"]["object_models"], fine_counts):
        cfg["use_fine"] = True
        cfg["positions_count_fine"] = pf
    return config, None, inputs


def scene_static_fine(seed=21):
    """Shipped field shape, 64 coarse + 64 fine samples per ray: the fine pass is a 128-sample tensor-core frame on explicit ray parameters."""
    config, _, inputs = with_fine(scene_static(seed=seed, P=64), [64])
    return config, scene_state(seed, config), inputs


def scene_tennis_fine(seed=22):
    """Tennis shape with fine models everywhere: court 4 + 4, players 32 + 32 samples per ray (ray benders on explicit ray parameters)."""
    config, _, inputs = with_fine(scene_tennis(seed=seed, height=64, width=64, stride=4, lead=(1, 1, 1), dense=True), [4, 32, 32])
    return config, scene_state(seed, config), inputs


def scene_toy_fine(seed=23, height=64, width=64, stride=4, lead=(1, 2, 1)):
    """Small well-conditioned networks (like toy_world, without the one-sample skybox -- the reference's sample_pdf needs >= 3 coarse
    samples): ground 8 + 8, a player model shared by 2 instances 12 + 12, fix_object_overlaps on; pins the fine pass's gradients."""
    toy = lambda: nerf_cfg(64, 4, 2, 4, 8)
    ground = object_cfg([[-10, 10], [-0.6, 2.0], [-10, 10]], 8, 0.05, 30.0, 16, 8, toy(), bender_cfg("zeroed"))
    player = object_cfg([[-0.6, 0.6], [0.0, 2.1], [-1.2, 1.2]], 12, 0.05, 30.0, 16, 8, toy(),
                        bender_cfg("positional", width=32, layers=3, skip=1, octaves=3))
    config = scene_config([ground, player], 1, [1, 2], True)
    for cfg, pf in zip(config["model"]["object_models"], [8, 12]):
        cfg["use_fine"] = True
        cfg["positions_count_fine"] = pf
    c2w = homogeneous(rot_x(-0.25), [0.0, 1.6, 6.0])
    orig, dirs, norm = camera_rays(lead, height, width, 0.8 * width, c2w, stride)
    pa = np.linalg.inv(homogeneous(rot_z(0.1), [-0.8, 0.0, 1.0]))
    pb = np.linalg.inv(homogeneous(np.eye(3), [1.0, 0.0, -0.5]))
    inputs = build_inputs(seed, config, lead, orig, dirs, norm, [np.eye(4), pa, pb])
    state = scene_state(seed, config)
    for k in list(state):
        if k.endswith("ray_bender.output_head.weight"):
            state[k] = state[k] * 8.0
    return config, state, inputs


FINE_SCENES = {
    "static_fine": lambda: scene_static_fine(),
    "tennis_fine": lambda: scene_tennis_fine(),
    "toy_fine": lambda: scene_toy_fine(),
}

SCENES = {
    "cfg1": lambda: scene_cfg1(),
    "static_small": lambda: scene_static(),
    "tennis_small": lambda: scene_tennis(),
    "tennis_dense": lambda: scene_tennis(seed=15, height=64, width=64, stride=4, lead=(1, 1, 1), dense=True),
    "tennis_anneal": lambda: scene_tennis(seed=16, height=64, width=64, stride=4, lead=(1, 1, 1), step=21000, dense=True),
    "minecraft_small": lambda: scene_minecraft(),
    "minecraft_absent": lambda: scene_minecraft(seed=17, absent=[(0, 0, 0, 3)]),
    "toy_world": lambda: scene_toy_world(),
}


# ----------------------------------------------------------------------------
# gradient pins: seeded cotangents and sub-sampling of large parameter gradients
# ----------------------------------------------------------------------------

GRAD_OUTPUT_KEYS = ("integrated_features", "opacity", "depth", "weights", "integrated_displacements_magnitude", "disparity")
GRAD_INPUT_KEYS = ("ray_origins", "ray_directions", "transformation_matrix_w2o", "style", "deformation")
GRAD_SAMPLE = 2048          # entries kept of parameter gradients larger than 2 * GRAD_SAMPLE


def _key_seed(key: str) -> int:
    import zlib
    return zlib.crc32(key.encode())


def cotangent(key: str, shape) -> torch.Tensor:
    """Seeded upstream gradient of output ``key`` ("object_0/opacity", "global/integrated_features", ...)."""
    rng = np.random.default_rng(_key_seed("cot:" + key))
    return torch.from_numpy(rng.normal(0.0, 1.0, tuple(shape)).astype(np.float32))


def grad_loss(coarse: dict, keys) -> torch.Tensor:
    """Scalar whose gradient is pinned: sum over ``keys`` ("<object>/<output>") of <cotangent, output>."""
    total = None
    for key in keys:
        obj, out = key.split("/")
        v = coarse[obj][out]
        term = (cotangent(key, v.shape).to(v.device) * v).sum()
        total = term if total is None else total + term
    return total


def grad_subsample(name: str, g: np.ndarray) -> np.ndarray:
    """Large gradients are stored as GRAD_SAMPLE seeded entries followed by [sum, L2 norm]."""
    flat = np.asarray(g, dtype=np.float32).reshape(-1)
    if flat.size <= 2 * GRAD_SAMPLE:
        return flat
    idx = np.random.default_rng(_key_seed("idx:" + name)).choice(flat.size, GRAD_SAMPLE, replace=False)
    return np.concatenate([flat[idx], np.array([flat.astype(np.float64).sum(), np.sqrt((flat.astype(np.float64) ** 2).sum())], dtype=np.float32)])
