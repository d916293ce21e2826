"""Golden outputs of the UPSTREAM ``ObjectComposer.forward_expected_positions`` (model/object_composer.py:624-722) on seeded scenes.

Run in the build container only:   python tests/golden/make_golden_expected.py
``expected_cases`` (shared with tests/test_gpu_parity.py) lists (scene, object id); the per-object inputs are slices of the scene's.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

EXPECTED_CASES = [("tennis_dense", 0), ("tennis_dense", 1), ("tennis_small", 2), ("minecraft_small", 0), ("minecraft_small", 2)]


def object_inputs(inputs, k):
    """The arguments of forward_expected_positions for object instance k of a scene (everything but object_id / perturb)."""
    return [inputs["ray_origins"], inputs["ray_directions"], inputs["focal_normals"], inputs["transformation_matrix_w2o"][..., k],
            inputs["style"][..., k], inputs["deformation"][..., k], inputs["object_in_scene"][..., k]]


def main():
    import make_golden as G           # applies the shims and imports the upstream composer
    import scenes
    out = {}
    for name, k in EXPECTED_CASES:
        config, state, inputs = scenes.SCENES[name]()
        comp = G.build_reference(config, state).eval()
        with torch.no_grad():
            exp, opacity = comp.forward_expected_positions(*object_inputs(inputs, k), k, False)["coarse"]
        out[f"{name}/{k}/expected_positions"] = exp.numpy()
        out[f"{name}/{k}/opacity"] = opacity.numpy()
        print(name, k, tuple(exp.shape), float(opacity.mean()))
    np.savez_compressed(os.path.join(HERE, "expected_positions.npz"), **out)


if __name__ == "__main__":
    main()
