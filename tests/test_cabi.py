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
