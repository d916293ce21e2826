"""Generate golden outputs by running the UPSTREAM reference code on the seeded scenes.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``model.object_composer.ObjectComposer`` from the reference tree with the
harness-side shims of SURVEY.md section 8c (never editing the reference), loads the seeded
parameters of ``scenes.py`` into it and stores the outputs of ``forward`` as
``tests/golden/<scene>[_variant].npz``.  Nothing at test/bench time imports the reference.
"""
from __future__ import annotations

import collections
import collections.abc
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REFERENCE = os.environ.get("PE_REFERENCE", "/root/reference")

# ---- shims ------------------------------------------------------
collections.Sequence = collections.abc.Sequence                      # ray_helper.py:217 etc. (py>=3.10)
np.bool = bool                                                       # object_composer.py:350 (numpy>=1.24)
torch.Tensor.cuda = lambda self, *a, **k: self                       # hard-coded .cuda() in ray_helper.py
_orig_randn_like = torch.randn_like
torch.randn_like = lambda t, **k: _orig_randn_like(t)                # object_composer.py:597 passes device=-1 on CPU
sys.path.insert(0, REFERENCE)

from model.object_composer import ObjectComposer  # noqa: E402  (upstream)
import scenes  # noqa: E402


def flatten(results: dict, prefix="") -> dict:
    out = {}
    for k, v in results.items():
        if torch.is_tensor(v):
            out[prefix + k] = v.detach().numpy()
        elif isinstance(v, dict):
            out.update(flatten(v, prefix + k + "/"))
    return out


class QueueRNG:
    """Feeds pre-generated tensors to the reference's torch.rand / torch.randn calls."""

    def __init__(self, rand, noise_list):
        self.rand, self.noise = list(rand), list(noise_list)

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn

        def rand(size, *a, **k):
            t = self.rand.pop(0)
            assert tuple(t.shape) == tuple(size), (t.shape, size)
            return t

        def randn(size, *a, **k):
            t = self.noise.pop(0)
            assert tuple(t.shape) == tuple(size), (t.shape, size)
            return t

        torch.rand, torch.randn = rand, randn
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn
        assert not self.rand and not self.noise, "unused perturbation tensors"


def build_reference(config, state):
    comp = ObjectComposer(copy.deepcopy(config))
    missing, unexpected = comp.load_state_dict(state, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    return comp


def run(name, variant="eval"):
    config, state, inputs = scenes.SCENES[name]()
    comp = build_reference(config, state)
    args = [inputs[k] for k in ("ray_origins", "ray_directions", "focal_normals", "transformation_matrix_w2o",
                                "style", "deformation", "object_in_scene")]
    extra = {}
    if variant == "eval":
        comp.eval()
        with torch.no_grad():
            res = comp(*args, False)
    elif variant == "perturb":
        comp.eval()
        rand, noise = scenes.perturbation_tensors(7, config, inputs)
        objs = len(rand)
        # call order in the reference: per object [rand (positions), randn (forward_object alphas)],
        # then per object randn (integrate), then the global randn (integrate of the composition)
        rand_q = list(rand)
        dummy = [torch.zeros_like(noise[f"object_{k}"]) for k in range(objs)]   # forward_object weights are unused (no fine model)
        noise_q = dummy + [noise[f"object_{k}"] for k in range(objs)] + [noise["global"]]
        with torch.no_grad(), QueueRNG(rand_q, noise_q):
            res = comp(*args, True)
    elif variant == "train":
        comp.train()
        # in the real pipeline positions depend on learnable camera/pose parameters; the Hutchinson
        # divergence (object_composer.py:597-598) differentiates w.r.t. them, so they must require grad
        args[1] = args[1].clone().requires_grad_(True)
        res = comp(*args, False)
        for k, v in comp.state_dict().items():
            if "running_" in k:
                extra["state/" + k] = v.detach().numpy()
    elif variant in ("grad", "grad_train"):
        run_grad(name, variant, comp, args)
        return
    else:
        raise ValueError(variant)
    flat = flatten(res)
    flat.update(extra)
    suffix = "" if variant == "eval" else "_" + variant
    path = os.path.join(HERE, f"{name}{suffix}.npz")
    np.savez_compressed(path, **flat)
    g = flat["coarse/global/integrated_features"]
    print(f"{name:18s} {variant:8s} -> {os.path.basename(path)}  keys={len(flat)}  "
          f"feat|mean|={np.abs(g).mean():.4f} opacity={flat['coarse/global/opacity'].mean():.4f} "
          f"size={os.path.getsize(path) / 1e3:.0f} kB")


def run_grad(name, variant, comp, args):
    """Gradients of the seeded scalar ``scenes.grad_loss`` w.r.t. every parameter and every differentiable input, computed by
    the upstream code's own autograd graph (eval-mode or train-mode BatchNorm)."""
    comp.train(variant == "grad_train")
    names = ("ray_origins", "ray_directions", "focal_normals", "transformation_matrix_w2o", "style", "deformation", "object_in_scene")
    leaves = {}
    for i, n in enumerate(names):
        if n in scenes.GRAD_INPUT_KEYS:
            args[i] = args[i].clone().requires_grad_(True)
            leaves[n] = args[i]
    res = comp(*args, False)["coarse"]
    keys = []
    for obj in res:
        for out in scenes.GRAD_OUTPUT_KEYS:
            v = res[obj][out]
            if out == "disparity" and not (torch.isfinite(v).all() and res[obj]["opacity"].min() > 0.5):
                continue            # 0/0 rays would poison every gradient with NaN
            if not v.requires_grad:
                continue
            keys.append(f"{obj}/{out}")
    loss = scenes.grad_loss(res, keys)
    loss.backward()
    flat = {"loss": np.array(loss.item(), dtype=np.float64), "loss_keys": np.array(keys)}
    for n, t in leaves.items():
        flat["input/" + n] = t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape), np.float32)
    for n, p in comp.named_parameters():
        g = p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
        flat["param/" + n] = scenes.grad_subsample(n, g)
    path = os.path.join(HERE, f"{name}_{variant}.npz")
    np.savez_compressed(path, **flat)
    gn = {k: float(np.abs(v).max()) for k, v in flat.items() if k.startswith("input/")}
    print(f"{name:18s} {variant:10s} -> {os.path.basename(path)} loss={loss.item():.4f} keys={len(keys)} "
          f"size={os.path.getsize(path) / 1e3:.0f} kB  max|input grads|={gn}")


if __name__ == "__main__":
    torch.manual_seed(0)
    for scene in scenes.SCENES:
        run(scene, "eval")
    run("cfg1", "perturb")
    run("tennis_dense", "perturb")
    run("minecraft_small", "perturb")
    run("cfg1", "train")
    run("static_small", "train")
    run("tennis_dense", "train")
    if "--no-grad" not in sys.argv:
        for scene in ("cfg1", "static_small", "tennis_dense", "minecraft_small", "toy_world"):
            run(scene, "grad")
        for scene in ("cfg1", "tennis_dense", "toy_world"):
            run(scene, "grad_train")
