"""Small driver for ncu (--profile-from-start off): ONE T-frame (bench.t_frame_report's scene, eval, mixed) after two warm-up frames."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import bench  # noqa: E402
import scenes  # noqa: E402
from helpers import INPUT_KEYS  # noqa: E402
from gpu_common import build_composer  # noqa: E402
from playableenvironments_b200.utils.lib_3d.ray_helper import RayHelper  # noqa: E402

torch.cuda.set_device(0)
H, W, strides = 288, 512, [4, 8]
lead = (1, 1, 1)
focal = 1700.0 * 0.51417 * 0.5 * (W / 256.0)
scene = bench.tennis_four_objects(lead, lambda st: scenes.camera_rays(lead, H, W, focal, scenes.tennis_camera(), st))
_, _, _, comp, dev = build_composer(scene, sys.argv[1] if len(sys.argv) > 1 else "mixed")
call = [dev[k] for k in INPUT_KEYS]
for i in range(3):
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.profiler.start()
    with torch.no_grad():
        if os.environ.get("PE_PROFILE_FOLD") == "1":
            feats = comp(*call, False)["coarse"]["global"]["integrated_features"]
            grids = RayHelper.fold_feature_grids(feats, strides, (H, W), [64, 128])
        else:
            grids = comp(*call, False, handoff=(strides, (H, W), [64, 128]))["coarse"]["global"]["feature_grids"]
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.profiler.stop()
print("done")
