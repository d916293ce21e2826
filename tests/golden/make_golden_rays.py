"""Golden outputs of the UPSTREAM ray-selection helpers (utils/lib_3d/ray_helper.py:55-183, 236-431, 611-795).

Run in the build container only:   python tests/golden/make_golden_rays.py
Inputs are seeded (``ray_cases`` below, shared with tests/test_ray_selection.py); the reference consumes torch's global CPU
generator (``torch.manual_seed(case seed)``), and so does the B200-side mirror, in the same order.
"""
from __future__ import annotations

import collections
import collections.abc
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("PE_REFERENCE", "/root/reference")


def ray_cases():
    """name -> (function name, seed, kwargs builder).  Inputs come from numpy's default_rng so only outputs are stored."""
    def base(seed, lead, h, w, objects):
        rng = np.random.default_rng(seed)
        dirs = torch.from_numpy(rng.standard_normal(lead + [h, w, 3]).astype(np.float32))
        obs = torch.from_numpy(rng.random(lead + [3, h, w]).astype(np.float32))
        lo = rng.random(lead + [2, objects]) * 0.6
        size = rng.random(lead + [2, objects]) * 0.35 + 0.02
        boxes = np.concatenate([lo, np.minimum(lo + size, 1.0)], axis=-2).astype(np.float32)      # (left, top, right, bottom)
        return dirs, obs, torch.from_numpy(boxes)

    cases = {}
    d, o, b = base(1, [2, 3, 1], 48, 64, 3)
    cases["strided_patch_2strides"] = ("sample_rays_strided_patch", 11, dict(ray_directions=d, observations=o, patch_size=8, strides=[4, 8],
                                                                            bounding_boxes=b, weights=[1.0, 2.0, 0.5], align_grid=True))
    d, o, b = base(2, [4, 1], 72, 128, 4)
    cases["strided_patch_tennis_like"] = ("sample_rays_strided_patch", 12, dict(ray_directions=d, observations=o, patch_size=16, strides=[4, 8],
                                                                               bounding_boxes=b, weights=[1.0, 1.0, 3.0, 3.0], align_grid=True))
    d, o, b = base(3, [3], 40, 40, 2)
    cases["strided_patch_single_stride"] = ("sample_rays_strided_patch", 13, dict(ray_directions=d, observations=o, patch_size=4, strides=2,
                                                                                 bounding_boxes=b, weights=[1.0, 1.0], align_grid=True))
    d, o, b = base(4, [2, 2], 32, 48, 3)
    b[0, 0, :, 1] = torch.tensor([0.5, 0.5, 0.5, 0.5])         # a zero-area box (guarded in sample_rays_weighted :672)
    cases["weighted"] = ("sample_rays_weighted", 14, dict(ray_directions=d, observations=o, samples_per_image=200, bounding_boxes=b, weights=[1.0, 2.0, 4.0]))
    cases["weighted_all"] = ("sample_rays_weighted", 15, dict(ray_directions=d, observations=o, samples_per_image=0, bounding_boxes=b, weights=[1.0, 2.0, 4.0]))
    cases["uniform"] = ("sample_rays", 16, dict(ray_directions=d, observations=o, samples_per_image=77))
    cases["uniform_all"] = ("sample_rays", 17, dict(ray_directions=d, observations=o, samples_per_image=0))
    d, o, b = base(5, [3, 1], 36, 52, 2)
    cases["patched"] = ("sample_rays_patched", 18, dict(ray_directions=d, observations=o, patch_size=6, patch_count=3, bounding_boxes=b, weights=[1.0, 3.0]))
    return cases


def fine_case():
    """Inputs of create_ray_positions_weighted (reference :1320-1347): coarse t values and weights of 3 x 50 rays x 16 samples."""
    rng = np.random.default_rng(21)
    lead, R, P = [3], 50, 16
    origins = torch.from_numpy(rng.standard_normal(lead + [3]).astype(np.float32))
    dirs = torch.from_numpy(rng.standard_normal(lead + [R, 3]).astype(np.float32))
    t = torch.from_numpy(np.sort(rng.random(lead + [R, P]).astype(np.float32) * 5.0 + 0.5, axis=-1))
    w = torch.from_numpy((rng.random(lead + [R, P]) ** 4).astype(np.float32))
    return origins, dirs, t, w


def main():
    collections.Sequence = collections.abc.Sequence
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REFERENCE)
    from utils.lib_3d.ray_helper import RayHelper as Ref  # noqa: E402  (upstream)

    out = {}
    for name, (fn, seed, kwargs) in ray_cases().items():
        torch.manual_seed(seed)
        res = getattr(Ref, fn)(**kwargs)
        for i, t in enumerate(res):
            out[f"{name}/{i}"] = t.numpy()
        print(name, [tuple(t.shape) for t in res])
    origins, dirs, t, w = fine_case()
    for perturb in (False, True):
        torch.manual_seed(33)
        pos, merged = Ref.create_ray_positions_weighted(origins, dirs, 24, t, w.clone(), perturb)     # (the reference adds 1e-5 to w in place)
        out[f"fine_{int(perturb)}/0"], out[f"fine_{int(perturb)}/1"] = pos.numpy(), merged.numpy()
        print("fine", perturb, tuple(pos.shape))
    np.savez_compressed(os.path.join(HERE, "ray_selection.npz"), **out)


if __name__ == "__main__":
    main()
