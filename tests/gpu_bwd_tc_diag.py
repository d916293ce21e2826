"""GPU diagnostic of the tensor-core field backward: gradients of the golden scenes with PE_BWD_TC=1 against the reference goldens AND
against the exact fp32 backward (PE_BWD_TC=0) of the same build, per tensor.  Usage: python tests/gpu_bwd_tc_diag.py [scene ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]

import scenes  # noqa: E402
from helpers import compare_grads  # noqa: E402
from test_gpu_backward import run_backward  # noqa: E402

CASES = [("static_small", False), ("minecraft_small", False), ("tennis_dense", False), ("tennis_dense", True)]


def rel(a, b):
    d = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max() / d)


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, training in CASES:
        if only and name not in only:
            continue
        os.environ["PE_BWD_TC"] = "0"
        golden, loss0, in0, par0 = run_backward(name, training)
        os.environ["PE_BWD_TC"] = "1"
        try:
            _, loss1, in1, par1 = run_backward(name, training)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"case": name, "training": training, "error": repr(e)[:500]}), flush=True)
            continue
        vs_gold = compare_grads(in1, par1, golden, -1.0)
        vs_gold0 = compare_grads(in0, par0, golden, -1.0)
        vs_fp32 = {"input/" + k: rel(in1[k], in0[k]) for k in in0}
        vs_fp32.update({"param/" + k: rel(par1[k], par0[k]) for k in par0})
        top = sorted(vs_fp32.items(), key=lambda kv: -(kv[1] if kv[1] == kv[1] else 1e30))[:12]
        print(json.dumps({"case": name, "training": training, "worst_vs_golden_tc": max(vs_gold.values()), "worst_vs_golden_fp32": max(vs_gold0.values()),
                          "worst_vs_fp32": [(k.replace("object_models_coarse.", "om."), float("%.3g" % v)) for k, v in top]}), flush=True)
