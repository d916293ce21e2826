"""Diagnostic: the T-train block of bench.py on its own (the Tennis training batch, this repo and the upstream composer in eager mode)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import bench
torch.cuda.set_device(0)
print(json.dumps(bench.t_train_report(torch.device("cuda", 0))))
