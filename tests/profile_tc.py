"""Small driver for ncu: runs the cfg2 workload (optionally reduced) a few times. Usage: profile_tc.py [side] [precision] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import scenes  # noqa: E402
from gpu_common import build_composer, run_composer  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scene = scenes.scene_static(seed=12, height=side, width=side, P=128)
_, _, _, comp, dev = build_composer(scene, precision)
for _ in range(reps):
    run_composer(comp, dev)
torch.cuda.synchronize()
print("done", side, precision)
