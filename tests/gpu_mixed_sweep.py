"""GPU diagnostic: the "mixed" precision mode (per-layer choice of hi / hi+lo weight passes, PE_TC_PASS2_MASK) against every golden
scene of the upstream reference and the 4096-ray golden of the full-size headline frame, with the kernel time of that frame.
Usage (GPU box): python tests/gpu_mixed_sweep.py [mask ...]   -> one JSON line per mask"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]

import scenes  # noqa: E402
from helpers import flatten, load_golden, scale_rel_err  # noqa: E402
from gpu_common import build_composer, run_composer  # noqa: E402

SCENES = ["cfg1", "static_small", "tennis_small", "tennis_dense", "tennis_anneal", "minecraft_small", "minecraft_absent", "toy_world"]


def golden_errors(precision):
    out = {}
    for name in SCENES:
        _, _, _, comp, dev = build_composer(name, precision)
        flat = flatten(run_composer(comp, dev))
        golden = load_golden(name)
        worst = (0.0, "")
        for k, ref in golden.items():
            if k.startswith("coarse/") and k in flat:
                worst = max(worst, (scale_rel_err(flat[k], ref), k))
        out[name] = [float("%.3g" % worst[0]), worst[1].replace("coarse/", "")]
    return out


def full_frame(precision, reps=5):
    scene = scenes.scene_static(seed=12, height=256, width=256, P=128)
    _, _, _, comp, dev = build_composer(scene, precision)
    run_composer(comp, dev)
    torch.cuda.synchronize()
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps):
        res = run_composer(comp, dev)["coarse"]["global"]
    e0.record()
    torch.cuda.synchronize()
    ms = s0.elapsed_time(e0) / reps
    g = load_golden("cfg2_subset")
    stride = int(g["stride"])
    out = {"ms": round(ms, 3), "frac_of_1637": round(8388608 * 1228288 / (ms * 1e-3) / 1e12 / 1637.0, 4)}
    for thr in (0.0, 4e-3):
        stable = np.abs(g["raw_alpha_last"].reshape(-1)) > thr
        for key in ("integrated_features", "opacity", "depth"):
            got = res[key].reshape(65536, -1)[::stride].cpu().numpy()[stable]
            want = g[key].reshape(4096, -1)[stable]
            out[f"{key}@{thr:g}"] = float("%.3g" % scale_rel_err(got, want))
        out[f"excluded@{thr:g}"] = float(1.0 - stable.mean())
    return out


if __name__ == "__main__":
    masks = sys.argv[1:] or ["fp16", "0x3F0", "0x3C0", "0x380", "0x300", "0x2F0", "fp16x2", "fp16x3"]
    for m in masks:
        if m.startswith("fp"):
            os.environ.pop("PE_TC_PASS2_MASK", None)
            precision = m
        else:
            os.environ["PE_TC_PASS2_MASK"] = m
            precision = "mixed"
        line = {"mode": m, "frame": full_frame(precision), "golden": golden_errors(precision)}
        print(json.dumps(line), flush=True)
