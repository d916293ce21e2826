"""Small driver for ncu (--profile-from-start off): ONE T-train step (bench.t_train_scene, forward + backward) after two warm-up steps.
Usage: profile_t_train.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import bench  # noqa: E402
import scenes  # noqa: E402
from helpers import INPUT_KEYS  # noqa: E402
from gpu_common import build_composer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.cuda.set_device(0)
scene, lead = bench.t_train_scene(B)
config, state, inputs, comp, dev = build_composer(scene, "mixed", training=True)
comp.allow_forward_without_grad = False
dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
call = [dev[k] for k in INPUT_KEYS]
rays = dev["ray_directions"].size(-2)
cot = torch.randn(lead + (rays, 192), device="cuda")
for step in range(3):
    comp.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    if step == 2:
        torch.cuda.profiler.start()
    res = comp(*call, True)["coarse"]
    loss = (res["global"]["integrated_features"] * cot).sum() + res["global"]["opacity"].sum()
    loss.backward()
    torch.cuda.synchronize()
    if step == 2:
        torch.cuda.profiler.stop()
print("done", B)
