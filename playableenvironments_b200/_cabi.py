"""ctypes binding of ``libpe_b200.so`` (C ABI declared in ``include/pe_b200.h``).

The library is built in-tree by ``csrc/build.sh`` (``__graft_entry__.build()``); importing this
module fails loudly when it is missing — there is no CPU or PyTorch fallback for the render path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

PE_ABI_VERSION = 9
PE_MAX_OBJECTS = 8
PE_MAX_LAYERS = 12
PE_MAX_OCTAVES = 16
PE_MAX_PEERS = 8

NERF_ADAIN, NERF_SKYBOX_V3 = 0, 1
BENDER_ZEROED, BENDER_POSITIONAL = 0, 1
PRECISION_FP32, PRECISION_FP16, PRECISION_FP16X2, PRECISION_FP16X3, PRECISION_MIXED = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": PRECISION_FP32, "fp16": PRECISION_FP16, "fp16x2": PRECISION_FP16X2, "fp16x3": PRECISION_FP16X3,
              "mixed": PRECISION_MIXED}

_f = C.POINTER(C.c_float)
_u8 = C.POINTER(C.c_uint8)


class PeObjectDesc(C.Structure):
    _fields_ = [
        ("nerf_kind", C.c_int32), ("bender_kind", C.c_int32),
        ("width", C.c_int32), ("layers", C.c_int32), ("skip", C.c_int32), ("octaves", C.c_int32), ("features", C.c_int32),
        ("style_features", C.c_int32), ("deformation_features", C.c_int32),
        ("b_width", C.c_int32), ("b_layers", C.c_int32), ("b_skip", C.c_int32), ("b_octaves", C.c_int32),
        ("positions", C.c_int32), ("is_static", C.c_int32), ("canonical_pose", C.c_int32),
        ("bbox", C.c_float * 6),
        ("z_near_min", C.c_float), ("z_far_max", C.c_float), ("empty_space_alpha", C.c_float),
        ("b_anneal", C.c_float * PE_MAX_OCTAVES),
        ("packed", C.c_void_p), ("aware_rounding", C.c_int32),
    ]


class PeObjectParams(C.Structure):
    _fields_ = [
        ("backbone_w", C.c_void_p * PE_MAX_LAYERS), ("backbone_b", C.c_void_p * PE_MAX_LAYERS),
        ("alpha_w", C.c_void_p), ("alpha_b", C.c_void_p),
        ("head0_w", C.c_void_p),
        ("affine1_w", C.c_void_p), ("affine1_b", C.c_void_p), ("bn1_mean", C.c_void_p), ("bn1_var", C.c_void_p),
        ("head3_w", C.c_void_p),
        ("affine2_w", C.c_void_p), ("affine2_b", C.c_void_p), ("bn2_mean", C.c_void_p), ("bn2_var", C.c_void_p),
        ("head6_w", C.c_void_p), ("head6_b", C.c_void_p),
        ("bender_w", C.c_void_p * PE_MAX_LAYERS), ("bender_b", C.c_void_p * PE_MAX_LAYERS),
        ("bender_out_w", C.c_void_p),
        ("backbone_in_moments", C.c_void_p * PE_MAX_LAYERS), ("head0_in_moments", C.c_void_p),
    ]


class PeScene(C.Structure):
    _fields_ = [
        ("images", C.c_int32), ("rays", C.c_int32), ("objects", C.c_int32), ("static_objects", C.c_int32),
        ("perturb", C.c_int32), ("training", C.c_int32), ("fix_object_overlaps", C.c_int32), ("apply_activation", C.c_int32),
        ("precision", C.c_int32), ("explicit_positions", C.c_int32), ("keep_samples", C.c_int32),
        ("explicit_t", C.c_int32), ("divergence", C.c_int32), ("bent_gradients", C.c_int32),
        ("object", PeObjectDesc * PE_MAX_OBJECTS),
        ("bwd_tiles", C.c_int64 * PE_MAX_OBJECTS),
    ]


class PeInputs(C.Structure):
    _fields_ = [
        ("ray_origins", C.c_void_p), ("ray_directions", C.c_void_p), ("w2o", C.c_void_p),
        ("style", C.c_void_p * PE_MAX_OBJECTS), ("deformation", C.c_void_p * PE_MAX_OBJECTS),
        ("object_in_scene", C.c_void_p),
        ("rand", C.c_void_p * PE_MAX_OBJECTS), ("noise", C.c_void_p * PE_MAX_OBJECTS), ("noise_global", C.c_void_p),
        ("positions", C.c_void_p), ("sample_t", C.c_void_p * PE_MAX_OBJECTS),
        ("divergence_noise", C.c_void_p * PE_MAX_OBJECTS), ("divergence_params", C.c_void_p),
    ]


class PeIntegrated(C.Structure):
    _fields_ = [
        ("integrated_features", C.c_void_p), ("opacity", C.c_void_p), ("weights", C.c_void_p), ("depth", C.c_void_p),
        ("disparity", C.c_void_p), ("integrated_displacements_magnitude", C.c_void_p), ("integrated_divergence", C.c_void_p),
    ]


PE_MAX_HANDOFF = 4


class PeHandoff(C.Structure):
    _fields_ = [
        ("segments", C.c_int32), ("ray_begin", C.c_int32 * PE_MAX_HANDOFF), ("ray_count", C.c_int32 * PE_MAX_HANDOFF),
        ("channel_begin", C.c_int32 * PE_MAX_HANDOFF), ("channel_count", C.c_int32 * PE_MAX_HANDOFF),
        ("grid", C.c_void_p * PE_MAX_HANDOFF),
    ]


class PeOutputs(C.Structure):
    _fields_ = [
        ("object", PeIntegrated * PE_MAX_OBJECTS), ("global_", PeIntegrated),
        ("raw_features", C.c_void_p * PE_MAX_OBJECTS), ("raw_alphas", C.c_void_p * PE_MAX_OBJECTS),
        ("displacements", C.c_void_p * PE_MAX_OBJECTS), ("positions_t", C.c_void_p * PE_MAX_OBJECTS),
        ("bn1_running", C.c_void_p * PE_MAX_OBJECTS), ("bn2_running", C.c_void_p * PE_MAX_OBJECTS),
        ("peers", C.c_int32), ("peer_features", C.c_void_p * PE_MAX_PEERS),
        ("handoff", PeHandoff),
    ]


class PeIntegratedGrads(C.Structure):
    _fields_ = [
        ("integrated_features", C.c_void_p), ("opacity", C.c_void_p), ("weights", C.c_void_p), ("depth", C.c_void_p),
        ("disparity", C.c_void_p), ("integrated_displacements_magnitude", C.c_void_p),
    ]


class PeOutGrads(C.Structure):
    _fields_ = [("object", PeIntegratedGrads * PE_MAX_OBJECTS), ("global_", PeIntegratedGrads),
                ("bent_positions", C.c_void_p * PE_MAX_OBJECTS)]


class PeObjectParamGrads(C.Structure):
    _fields_ = [
        ("backbone_w", C.c_void_p * PE_MAX_LAYERS), ("backbone_b", C.c_void_p * PE_MAX_LAYERS),
        ("alpha_w", C.c_void_p), ("alpha_b", C.c_void_p),
        ("head0_w", C.c_void_p),
        ("affine1_w", C.c_void_p), ("affine1_b", C.c_void_p),
        ("head3_w", C.c_void_p),
        ("affine2_w", C.c_void_p), ("affine2_b", C.c_void_p),
        ("head6_w", C.c_void_p), ("head6_b", C.c_void_p),
        ("bender_w", C.c_void_p * PE_MAX_LAYERS), ("bender_b", C.c_void_p * PE_MAX_LAYERS),
        ("bender_out_w", C.c_void_p),
    ]


class PeInGrads(C.Structure):
    _fields_ = [
        ("ray_origins", C.c_void_p), ("ray_directions", C.c_void_p), ("w2o", C.c_void_p),
        ("style", C.c_void_p * PE_MAX_OBJECTS), ("deformation", C.c_void_p * PE_MAX_OBJECTS),
        ("params", PeObjectParamGrads * PE_MAX_OBJECTS), ("sample_t", C.c_void_p * PE_MAX_OBJECTS),
    ]


EXPORTS = [
    "pe_abi_version", "pe_last_error", "pe_take_launch_count", "pe_packed_bytes", "pe_pack_object",
    "pe_workspace_bytes", "pe_render_forward", "pe_backward_workspace_bytes", "pe_render_backward", "pe_render_backward_saved", "pe_forward_tile_counts", "pe_positional_encoding", "pe_generate_rays",
    "pe_fold_feature_grids", "pe_debug_umma_gemm", "pe_debug_umma_gemm2", "pe_debug_pack_layer",
]

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpe_b200.so")
_lib: Optional[C.CDLL] = None


class PeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      f"(or playableenvironments_b200/csrc/build.sh). There is no fallback path.")
    L = C.CDLL(LIB_PATH)
    L.pe_abi_version.restype = C.c_int
    L.pe_last_error.restype = C.c_char_p
    L.pe_take_launch_count.restype = C.c_int64
    L.pe_packed_bytes.restype = C.c_size_t
    L.pe_packed_bytes.argtypes = [C.POINTER(PeObjectDesc)]
    L.pe_pack_object.restype = C.c_int
    L.pe_pack_object.argtypes = [C.POINTER(PeObjectDesc), C.POINTER(PeObjectParams), C.c_void_p, C.c_void_p]
    L.pe_workspace_bytes.restype = C.c_size_t
    L.pe_workspace_bytes.argtypes = [C.POINTER(PeScene)]
    L.pe_render_forward.restype = C.c_int
    L.pe_render_forward.argtypes = [C.POINTER(PeScene), C.POINTER(PeInputs), C.POINTER(PeOutputs), C.c_void_p, C.c_size_t, C.c_void_p]
    L.pe_backward_workspace_bytes.restype = C.c_size_t
    L.pe_backward_workspace_bytes.argtypes = [C.POINTER(PeScene)]
    L.pe_render_backward.restype = C.c_int
    L.pe_render_backward.argtypes = [C.POINTER(PeScene), C.POINTER(PeInputs), C.POINTER(PeObjectParams), C.POINTER(PeOutGrads),
                                     C.POINTER(PeInGrads), C.c_void_p, C.c_size_t, C.c_void_p]
    L.pe_render_backward_saved.restype = C.c_int
    L.pe_render_backward_saved.argtypes = [C.POINTER(PeScene), C.POINTER(PeInputs), C.POINTER(PeObjectParams), C.POINTER(PeOutGrads),
                                           C.POINTER(PeInGrads), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    L.pe_forward_tile_counts.restype = C.c_int
    L.pe_forward_tile_counts.argtypes = [C.POINTER(PeScene), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.pe_positional_encoding.restype = C.c_int
    L.pe_positional_encoding.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pe_generate_rays.restype = C.c_int
    L.pe_generate_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pe_fold_feature_grids.restype = C.c_int
    L.pe_fold_feature_grids.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p), C.c_void_p]
    L.pe_debug_umma_gemm.restype = C.c_int
    L.pe_debug_umma_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.pe_debug_pack_layer.restype = C.c_int
    L.pe_debug_pack_layer.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pe_debug_umma_gemm2.restype = C.c_int
    L.pe_debug_umma_gemm2.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    if L.pe_abi_version() != PE_ABI_VERSION:
        raise PeError(f"libpe_b200.so ABI {L.pe_abi_version()} != binding {PE_ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise PeError(f"pe_b200 error {rc}: {lib().pe_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise PeError("the B200 render path needs CUDA tensors; there is no CPU path")
    if not t.is_contiguous():
        raise PeError("internal: non-contiguous tensor handed to the C ABI")
    return t.data_ptr()


def current_stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()
