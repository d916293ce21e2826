"""Diagnostic: worst scale-relative error of every output of the golden scenes / the full-size frame per mode, with the activation-aware
weight stream on (default) and off (PE_TC_AWARE=0), plus step times of the full-size frame."""
import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE, os.path.join(_HERE, "golden")):
    sys.path.insert(0, _p)
import numpy as np, torch, scenes
from helpers import flatten, load_golden, scale_rel_err
from gpu_common import build_composer, run_composer

def worst(name, precision):
    _, _, _, comp, dev = build_composer(name, precision)
    got = flatten(run_composer(comp, dev)); torch.cuda.synchronize()
    g = load_golden(name)
    errs = {k: scale_rel_err(got[k], v) for k, v in g.items() if k.startswith("coarse/") and "divergence" not in k}
    k = max(errs, key=errs.get)
    return k.replace("coarse/", ""), errs[k]

def fullsize(precision):
    scene = scenes.scene_static(seed=12, height=256, width=256, P=128)
    _, _, _, comp, dev = build_composer(scene, precision)
    full = run_composer(comp, dev)["coarse"]["global"]
    g = load_golden("cfg2_subset"); stride = int(g["stride"])
    stable = np.abs(g["raw_alpha_last"].reshape(-1)) > 4e-3
    out = {}
    for key in ("integrated_features", "opacity", "depth"):
        got = full[key].reshape(65536, -1)[::stride].cpu().numpy()[stable]
        out[key] = scale_rel_err(got, g[key].reshape(4096, -1)[stable])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3): run_composer(comp, dev)
    ev[0].record()
    for _ in range(10): run_composer(comp, dev)
    ev[1].record(); torch.cuda.synchronize()
    return out, ev[0].elapsed_time(ev[1]) / 10

masks = sys.argv[1:] or ["0x000", "0x080", "0x0C0", "0x0E0", "0x100", "0x180"]
for mask in masks:
    os.environ["PE_TC_AWARE_MASK"] = mask
    w = worst('static_small', 'mixed')
    errs, ms = fullsize('mixed')
    print(f"aware mask {mask}: static_small worst {w[0]} {w[1]:.3e}; full-size {({k: float(f'{v:.3g}') for k, v in errs.items()})} {ms:.2f} ms/frame", flush=True)
