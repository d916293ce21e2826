"""cfg3 Tennis train step (forward + backward through ObjectComposer) timing: python tests/gpu_train_step.py [dense|sparse] [reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import bench  # noqa: E402

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    dev = torch.device("cuda", 0)
    for dense in ([True] if which == "dense" else [False] if which == "sparse" else [False, True]):
        print(json.dumps(bench.train_step_report(dev, dense)), flush=True)
