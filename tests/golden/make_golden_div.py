"""Goldens of (a) the Hutchinson divergence (model/object_composer.py:582-601) and (b) the gradients of
``forward_expected_positions`` (:603-722), produced by the UPSTREAM composer.  Build container only:

    python tests/golden/make_golden_div.py

(a) ``<scene>_div.npz``: a train-mode forward whose ``torch.randn_like`` probe vectors come from ``scenes.divergence_noise`` (one per
    object instance in call order -- the reference draws one for zeroed ray benders too when the positions carry a graph, and gets 0).
(b) ``expected_positions_grad.npz``: for (scene, object) cases the gradients of <c1, expected_positions> + <c2, opacity> w.r.t. every
    parameter of the object's model and the differentiable inputs."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import scenes  # noqa: E402
from make_golden_expected import object_inputs  # noqa: E402

ARG_KEYS = ("ray_origins", "ray_directions", "focal_normals", "transformation_matrix_w2o", "style", "deformation", "object_in_scene")
DIV_SCENES = ["toy_world", "tennis_dense"]
from make_golden_div_cases import EXPECTED_GRAD_CASES, expected_loss  # noqa: E402


def run_divergence(name):
    config, state, inputs = scenes.SCENES[name]()
    comp = MG.build_reference(config, state).train()
    args = [inputs[k] for k in ARG_KEYS]
    args[1] = args[1].clone().requires_grad_(True)         # positions must carry a graph (see make_golden.py, variant "train")
    queue = [e for e in scenes.divergence_noise(9, config, inputs)]
    orig = torch.randn_like

    def randn_like(t, **kw):
        e = queue.pop(0)
        assert tuple(e.shape) == tuple(t.shape), (e.shape, t.shape)
        return e

    torch.randn_like = randn_like
    try:
        res = comp(*args, False)
    finally:
        torch.randn_like = orig
    assert not queue
    flat = MG.flatten(res)
    path = os.path.join(HERE, f"{name}_div.npz")
    np.savez_compressed(path, **flat)
    d = flat["coarse/global/integrated_divergence"]
    print(f"{name:14s} div -> {os.path.basename(path)} mean integrated divergence {d.mean():.4e} max {d.max():.4e}")


def run_expected_grads():
    out = {}
    for name, k in EXPECTED_GRAD_CASES:
        config, state, inputs = scenes.SCENES[name]()
        comp = MG.build_reference(config, state).eval()
        args = object_inputs(inputs, k)
        leaves = {}
        for i, n in enumerate(ARG_KEYS):
            if n in scenes.GRAD_INPUT_KEYS:
                args[i] = args[i].clone().requires_grad_(True)
                leaves[n] = args[i]
        exp, opacity = comp.forward_expected_positions(*args, k, False)["coarse"]
        loss = expected_loss(name, k, exp, opacity)
        loss.backward()
        out[f"{name}/{k}/loss"] = np.array(loss.item(), dtype=np.float64)
        for n, t in leaves.items():
            out[f"{name}/{k}/input/{n}"] = t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape), np.float32)
        for n, p in comp.named_parameters():
            if p.grad is not None:
                out[f"{name}/{k}/param/{n}"] = scenes.grad_subsample(n, p.grad.numpy())
        print(f"{name:14s} object {k}: expected-positions gradients, loss {loss.item():.4f}, "
              f"{sum(1 for key in out if key.startswith(f'{name}/{k}/param/'))} parameter tensors")
    np.savez_compressed(os.path.join(HERE, "expected_positions_grad.npz"), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    for scene in DIV_SCENES:
        run_divergence(scene)
    run_expected_grads()
