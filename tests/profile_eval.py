"""Small driver for ncu: cfg3 (Tennis) EVAL forward through ObjectComposer. Usage: profile_eval.py [dense:0|1] [precision]"""
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
scene = scenes.scene_tennis(seed=13, height=144, width=256, stride=1, lead=(1, 1, 1), dense=dense)
_, _, _, comp, dev = build_composer(scene, precision)
call = [dev[k] for k in INPUT_KEYS]
with torch.no_grad():
    for _ in range(3):
        comp(*call, False)
torch.cuda.synchronize()
print("done", dense, precision)
