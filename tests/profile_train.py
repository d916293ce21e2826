"""Small driver for ncu: cfg3 (Tennis, train mode) forward + backward through ObjectComposer, one warm-up and one measured step.
Usage: profile_train.py [dense:0|1] [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import scenes  # noqa: E402
from helpers import INPUT_KEYS  # noqa: E402
from gpu_common import build_composer  # noqa: E402

dense = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
torch.cuda.set_device(0)
scene = scenes.scene_tennis(seed=13, height=144, width=256, stride=1, lead=(1, 1, 1), dense=dense)
config, state, inputs, comp, dev = build_composer(scene, precision, training=True)
comp.allow_forward_without_grad = False
dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
call = [dev[k] for k in INPUT_KEYS]
rays = dev["ray_directions"].size(-2)
cot = torch.randn(rays, 192, device="cuda")
for step in range(2):
    comp.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    print("step", step, flush=True)
    res = comp(*call, False)["coarse"]
    loss = (res["global"]["integrated_features"].reshape(rays, 192) * cot).sum() + res["global"]["opacity"].sum()
    loss.backward()
torch.cuda.synchronize()
print("done", dense, precision)
