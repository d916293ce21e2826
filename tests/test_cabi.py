"""The C-ABI library loads (no GPU needed) and exports every symbol include/pe_b200.h declares; the ctypes mirrors
of the POD structs have the same size as the C definitions."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pe_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pe_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from playableenvironments_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return _cabi.lib()


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for expected in ("pe_abi_version", "pe_render_forward", "pe_pack_object", "pe_workspace_bytes", "pe_positional_encoding",
                     "pe_generate_rays", "pe_fold_feature_grids"):
        assert expected in names


def test_library_exports_every_declared_symbol(lib):
    from playableenvironments_b200 import _cabi
    names = declared_functions()
    assert sorted(_cabi.EXPORTS) == names, "ctypes binding and header disagree"
    for name in names:
        assert hasattr(lib, name), f"libpe_b200.so does not export {name}"
    assert lib.pe_abi_version() == _cabi.PE_ABI_VERSION


def test_struct_sizes_match_the_c_definitions():
    from playableenvironments_b200 import _cabi
    src = '#include <stdio.h>\n#include "pe_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(PeObjectDesc), sizeof(PeObjectParams), ' \
          'sizeof(PeScene), sizeof(PeInputs), sizeof(PeIntegrated), sizeof(PeOutputs), sizeof(PeIntegratedGrads), sizeof(PeOutGrads), ' \
          'sizeof(PeObjectParamGrads), sizeof(PeInGrads));return 0;}\n'
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "sizes.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "sizes")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    mine = [ctypes.sizeof(t) for t in (_cabi.PeObjectDesc, _cabi.PeObjectParams, _cabi.PeScene, _cabi.PeInputs, _cabi.PeIntegrated, _cabi.PeOutputs,
                                       _cabi.PeIntegratedGrads, _cabi.PeOutGrads, _cabi.PeObjectParamGrads, _cabi.PeInGrads)]
    assert mine == sizes


def test_argument_validation_without_a_gpu(lib):
    """Host-side validation paths return an error code and a message; no compute call is made."""
    from playableenvironments_b200 import _cabi
    d = _cabi.PeObjectDesc()
    d.width, d.layers, d.skip, d.features, d.positions = 64, 4, 9, 3, 16          # skip >= layers
    assert lib.pe_packed_bytes(ctypes.byref(d)) == 0
    assert b"Skip layer" in lib.pe_last_error()
    d.skip = 2
    d.octaves = 4
    assert lib.pe_packed_bytes(ctypes.byref(d)) > 0
    scene = _cabi.PeScene()
    scene.objects = 0
    assert lib.pe_workspace_bytes(ctypes.byref(scene)) == 0


def test_cpu_tensors_fail_loudly():
    """There is no CPU fallback: handing CPU tensors to the render path raises."""
    import torch
    import scenes
    from helpers import INPUT_KEYS
    from playableenvironments_b200 import _cabi
    from playableenvironments_b200.model.object_composer import ObjectComposer
    config, state, inputs = scenes.SCENES["cfg1"]()
    comp = ObjectComposer(config)
    comp.load_state_dict(state, strict=False)
    comp.eval()
    with torch.no_grad(), pytest.raises(_cabi.PeError):
        comp(*[inputs[k] for k in INPUT_KEYS], False)


def _shipped_desc(positional_bender: bool, positions: int):
    from playableenvironments_b200 import _cabi
    d = _cabi.PeObjectDesc()
    d.nerf_kind, d.bender_kind = _cabi.NERF_ADAIN, (_cabi.BENDER_POSITIONAL if positional_bender else _cabi.BENDER_ZEROED)
    d.width, d.layers, d.skip, d.octaves, d.features, d.positions = 256, 8, 4, 10, 192, positions
    d.style_features, d.deformation_features = 64, 32
    d.b_width, d.b_layers, d.b_skip, d.b_octaves = 128, 6, 3, 6
    d.bbox[:] = [-1.0, 1.0, -1.0, 1.0, 0.0, 2.0]
    d.z_near_min, d.z_far_max, d.empty_space_alpha = 0.1, 10.0, -3.5
    d.packed = 256                                  # any non-null pointer: sizing never dereferences it
    return d


def test_workspace_and_blob_sizes_follow_the_chosen_path(lib):
    """Host-side sizing (no GPU): the packed blob of a shipped-shape object holds the tensor-core weight streams (field: hi + lo passes
    twice -- single-CTA and CTA-pair layout --, ray bender: hi + lo), and the forward workspace grows by the pre-pass hand-off (bent
    positions, masks, tile list) exactly when a ray-bender object takes the tensor-core path."""
    from playableenvironments_b200 import _cabi
    plain, bent = _shipped_desc(False, 32), _shipped_desc(True, 32)
    b_plain, b_bent = lib.pe_packed_bytes(ctypes.byref(plain)), lib.pe_packed_bytes(ctypes.byref(bent))
    assert b_plain > 4 * 1_300_000                  # 2 layouts x 2 passes x 1.30 MB of fp16 slabs (+ the fp32 section)
    assert b_bent - b_plain > 2 * 241_664           # + bender fp32 weights + its two slab passes
    scene = _cabi.PeScene()
    scene.images, scene.rays, scene.objects, scene.static_objects = 2, 1000, 1, 0
    scene.object[0] = bent
    sizes = {}
    for name in ("fp32", "fp16x3"):
        scene.precision = _cabi.PRECISIONS[name]
        sizes[name] = lib.pe_workspace_bytes(ctypes.byref(scene))
        assert sizes[name] > 0, lib.pe_last_error()
    slots = 2 * 1000 * 32
    assert sizes["fp16x3"] - sizes["fp32"] >= slots * 13            # 12 B position + 1 B mask per slot (+ tile list)
    scene.training = 1
    assert lib.pe_backward_workspace_bytes(ctypes.byref(scene)) > sizes["fp16x3"] + slots * 256 * 4      # + trunk-output hand-off


def test_debug_entry_points_validate_their_arguments(lib):
    assert lib.pe_debug_umma_gemm2(3, None, None, None, 128, 64, 0, 0, None) != 0
    assert b"debug gemm2" in lib.pe_last_error()
    assert lib.pe_debug_umma_gemm2(1, None, None, None, 100, 64, 128, 2048, None) != 0
