"""Goldens of the hierarchical ("fine") pass, produced by the UPSTREAM ObjectComposer with ``use_fine`` object models
(model/object_composer.py:44-53, 561-578; utils/lib_3d/ray_helper.py:1320-1403).  Build container only:

    python tests/golden/make_golden_fine.py

Writes ``<scene>.npz`` (eval forward: every ``coarse/...`` and ``fine/...`` output), ``toy_fine_train.npz`` (train-mode BatchNorm +
running statistics of the coarse AND fine models) and ``<scene>_grad.npz`` for toy_fine / static_fine (gradients of a seeded scalar
over the coarse and the fine results w.r.t. every parameter and differentiable input, by the upstream autograd graph)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the SURVEY 8c shims and imports the upstream composer)
import scenes  # noqa: E402

ARG_KEYS = ("ray_origins", "ray_directions", "focal_normals", "transformation_matrix_w2o", "style", "deformation", "object_in_scene")


def fine_loss(res, keys):
    """scenes.grad_loss over both result families; keys are "<coarse|fine>/<object>/<output>"."""
    total = None
    for family in ("coarse", "fine"):
        sub = [k.split("/", 1)[1] for k in keys if k.startswith(family + "/")]
        # cotangents are seeded by the full key so that the two families get different ones
        for key in sub:
            obj, out = key.split("/")
            v = res[family][obj][out]
            term = (scenes.cotangent(f"{family}/{key}", v.shape).to(v.device) * v).sum()
            total = term if total is None else total + term
    return total


def loss_keys(res):
    keys = []
    for family in ("coarse", "fine"):
        for obj in res[family]:
            for out in scenes.GRAD_OUTPUT_KEYS:
                v = res[family][obj][out]
                if out == "disparity" and not (torch.isfinite(v).all() and res[family][obj]["opacity"].min() > 0.5):
                    continue
                if v.requires_grad:
                    keys.append(f"{family}/{obj}/{out}")
    return keys


def run(name, variant):
    config, state, inputs = scenes.FINE_SCENES[name]()
    comp = MG.build_reference(config, state)
    args = [inputs[k] for k in ARG_KEYS]
    flat = {}
    if variant == "eval":
        comp.eval()
        with torch.no_grad():
            flat = MG.flatten(comp(*args, False))
    elif variant == "train":
        comp.train()
        args[1] = args[1].clone().requires_grad_(True)      # the Hutchinson term differentiates w.r.t. the positions (see make_golden.py)
        flat = MG.flatten(comp(*args, False))
        for k, v in comp.state_dict().items():
            if "running_" in k:
                flat["state/" + k] = v.detach().numpy()
    elif variant == "grad":
        comp.eval()
        leaves = {}
        for i, n in enumerate(ARG_KEYS):
            if n in scenes.GRAD_INPUT_KEYS:
                args[i] = args[i].clone().requires_grad_(True)
                leaves[n] = args[i]
        res = comp(*args, False)
        keys = loss_keys(res)
        loss = fine_loss(res, keys)
        loss.backward()
        flat = {"loss": np.array(loss.item(), dtype=np.float64), "loss_keys": np.array(keys)}
        for n, t in leaves.items():
            flat["input/" + n] = t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape), np.float32)
        for n, p in comp.named_parameters():
            g = p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
            flat["param/" + n] = scenes.grad_subsample(n, g)
    suffix = "" if variant == "eval" else "_" + variant
    path = os.path.join(HERE, f"{name}{suffix}.npz")
    np.savez_compressed(path, **flat)
    print(f"{name:12s} {variant:6s} -> {os.path.basename(path)} keys={len(flat)} size={os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    torch.manual_seed(0)
    for scene in scenes.FINE_SCENES:
        run(scene, "eval")
    run("toy_fine", "train")
    run("toy_fine", "grad")
    run("static_fine", "grad")
