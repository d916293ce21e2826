"""Diagnostic: golden-scene errors and Tennis frame times of the mixed mode when objects with few samples per ray are NOT sent to fp16x3
(PE_TC_X3_BELOW) and use the activation-aware stream with a given two-pass mask (PE_TC_AWARE_MIN_POSITIONS=0, PE_TC_AWARE_MASK)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
from gpu_mixed_sweep import golden_errors
import bench
for below, mask in (("64", "0x0C0"), ("0", "0x7FF"), ("0", "0x0F8"), ("0", "0x0C0"), ("5", "0x0F8"), ("5", "0x0C0")):
    os.environ["PE_TC_X3_BELOW"] = below
    os.environ["PE_TC_AWARE_MIN_POSITIONS"] = "0"
    os.environ["PE_TC_AWARE_MASK"] = mask
    g = golden_errors("mixed")
    ev = bench.eval_frame_report(torch.device("cuda", 0), True)
    print(json.dumps({"x3_below": below, "mask": mask, "golden": {k: v[0] for k, v in g.items()}, "worst_key": {k: v[1] for k, v in g.items() if v[0] > 1e-3},
                      "dense_eval_mixed_ms": round(ev["mixed_ms"], 3)}), flush=True)
