"""Small driver for ncu: one cfg3 (Tennis, train mode) forward+backward through ObjectComposer. Usage: profile_train.py [dense:0|1]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import bench  # noqa: E402

dense = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
torch.cuda.set_device(0)
print(bench.train_step_report(torch.device("cuda", 0), dense))
