"""Diagnostic (CPU, build container): which layers' fp16 weight rounding dominates the error of the single-pass tensor-core
mode?  Emulates the kernel's numerics inside the oracle (fp16 operands, fp32 accumulation, zero-sum weight rounding, head
layer 6 in fp32) and lets chosen layers run with hi+lo weights (two MMA passes).  Usage: emulate_precision.py [side]"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import scenes  # noqa: E402
from helpers import INPUT_KEYS  # noqa: E402
from oracle import render_oracle as O  # noqa: E402


def zero_sum_fp16(w: torch.Tensor) -> torch.Tensor:
    """pe_tc_pack_layer_kernel: walk along K, pick the fp16 neighbour that keeps the row's running rounding error near 0."""
    w = w.numpy()
    near = w.astype(np.float16).astype(np.float32)
    up = np.nextafter(near.astype(np.float16), np.float16(np.inf)).astype(np.float32)
    down = np.nextafter(near.astype(np.float16), np.float16(-np.inf)).astype(np.float32)
    other = np.where(near < w, up, np.where(near > w, down, near))
    out = np.empty_like(w)
    run = np.zeros(w.shape[0], dtype=np.float32)
    for k in range(w.shape[1]):
        e_near, e_other = near[:, k] - w[:, k], other[:, k] - w[:, k]
        pick = np.abs(run + e_other) < np.abs(run + e_near)
        out[:, k] = np.where(pick, other[:, k], near[:, k])
        run += np.where(pick, e_other, e_near)
    return torch.from_numpy(out)


def run(side: int):
    scene = scenes.scene_static(seed=12, height=side, width=side, P=128)
    config, state, inputs = scene
    args = [inputs[k] for k in INPUT_KEYS]
    ref_all = O.composer_forward(config, state, *args, perturb=False)["coarse"]["global"]
    ref, ref_op = ref_all["integrated_features"], ref_all["opacity"]
    orig_linear = O._linear
    cache = {}

    def make(two_pass: set):
        def lin(sd, prefix, x):
            name = prefix.split("object_models_coarse.0.")[-1]
            w, b = sd[prefix + ".weight"], sd.get(prefix + ".bias")
            if "affine_transform" in name or name.endswith("alpha_head") or name.endswith("features_head.6"):
                return F.linear(x, w, b)                      # fp32 in the kernel (style prologue, alpha head, folded head layer 6)
            if name not in cache:
                cache[name] = zero_sum_fp16(w)
            wq = w if name in two_pass else cache[name]       # hi + lo = 22 significant bits: exact for this purpose
            xq = x.half().float()
            return F.linear(xq, wq, b)
        return lin

    names = [f"nerf_model.backbone_layers.{i}" for i in range(8)] + ["nerf_model.features_head.0", "nerf_model.features_head.3"]
    variants = {"all single pass": set(), "L0 x2": {names[0]}, "H0,H3 x2": set(names[8:]), "L0,H0,H3 x2": {names[0], *names[8:]},
                "L7,H0,H3 x2": set(names[7:]), "L4-L7,H0,H3 x2": set(names[4:]), "all x2 (fp16x2)": set(names)}
    for label, tp in variants.items():
        O._linear = make(tp)
        try:
            res = O.composer_forward(config, state, *args, perturb=False)["coarse"]["global"]
        finally:
            O._linear = orig_linear
        # rays on the opacity step (last sample, interval 1e10: alpha flips 0 <-> 1 with the sign of its raw value) are excluded
        stable = ((res["opacity"] - ref_op).abs() < 5e-3).reshape(-1)
        d = (res["integrated_features"] - ref).double().reshape(-1, ref.size(-1))[stable]
        r = ref.double().reshape(-1, ref.size(-1))[stable]
        print(f"{label:22s} max/scale {float(d.abs().max() / r.abs().max()):.3e}  rel_l2 {float(d.norm() / r.norm()):.3e}  "
              f"step rays {float(1 - stable.float().mean()):.4f}")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        run(int(sys.argv[1]) if len(sys.argv) > 1 else 48)
