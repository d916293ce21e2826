"""Golden of the FULL-SIZE headline frame (BASELINE configs[1]: 256x256 rays x 128 samples/ray, bench.py's workload) on a strided
subset of its rays, produced by the UPSTREAM ObjectComposer (build container only):

    python tests/golden/make_golden_fullsize.py

Every 16th ray of the frame (4096 rays, 524 288 samples) -> tests/golden/cfg2_subset.npz.  Rays are independent, so the subset of the
full-frame render must equal the render of the subset.  Stored: the composed scene's integrated_features / opacity / depth and the raw
alpha of each ray's last sample (the reference's opacity is a step function of it: interval 1e10, object_composer.py:172,197).
"""
import os

import numpy as np
import torch

import make_golden as M          # shims + upstream import
import scenes

STRIDE = 16
KEYS = ("ray_origins", "ray_directions", "focal_normals", "transformation_matrix_w2o", "style", "deformation", "object_in_scene")


def main():
    torch.manual_seed(0)
    config, state, inputs = scenes.scene_static(seed=12, height=256, width=256, P=128)
    comp = M.build_reference(config, state)
    comp.eval()
    inputs = dict(inputs)
    inputs["ray_directions"] = inputs["ray_directions"][..., ::STRIDE, :].contiguous()
    rays = inputs["ray_directions"].size(-2)
    feats, opacity, depth, raw_last = [], [], [], []
    # raw alphas are not part of the composer's result: taken from the object model's own output (second return value)
    comp.object_models_coarse[0].register_forward_hook(lambda module, args, output: raw_last.append(output[1][..., -1].detach().clone()))
    with torch.no_grad():
        for begin in range(0, rays, 512):
            chunk = dict(inputs)
            chunk["ray_directions"] = inputs["ray_directions"][..., begin:begin + 512, :]
            res = comp(*[chunk[k] for k in KEYS], False)["coarse"]
            feats.append(res["global"]["integrated_features"])
            opacity.append(res["global"]["opacity"])
            depth.append(res["global"]["depth"])
    out = {"integrated_features": torch.cat(feats, -2).numpy(), "opacity": torch.cat(opacity, -1).numpy(),
           "depth": torch.cat(depth, -1).numpy(), "raw_alpha_last": torch.cat(raw_last, -1).numpy(), "stride": np.array(STRIDE)}
    path = os.path.join(M.HERE, "cfg2_subset.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()}, os.path.getsize(path) / 1e6, "MB", "has_raw", bool(np.abs(out["raw_alpha_last"]).max() > 0))


if __name__ == "__main__":
    main()
