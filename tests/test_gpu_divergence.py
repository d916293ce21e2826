"""(a) The Hutchinson divergence of the ray benders' displacement fields (model/object_composer.py:582-601) and (b) the backward of
``forward_expected_positions`` (:603-722) on the GPU, against outputs / gradients of the upstream composer
(tests/golden/make_golden_div.py).  Both ride on the same device pass: a vector-Jacobian product through the ray bender (the
ray-bender-only mode of pe_field_bwd_kernel) with a caller-supplied upstream vector per sample."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE, os.path.join(_HERE, "golden")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import pytest
import torch

import scenes
from helpers import INPUT_KEYS, compare, flatten, load_golden, scale_rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,precision,tol", [("toy_world", "fp32", 2e-4), ("tennis_dense", "fp32", 3e-4), ("tennis_dense", "fp16x3", 5e-4),
                                                ("tennis_dense", "mixed", 2e-3)])
def test_hutchinson_divergence_matches_reference(name, precision, tol):
    """Train-mode forward with ``compute_divergence`` and the probe vectors of the golden run: integrated_divergence of every object and of
    the composed scene (mean of alpha |e . J e|), next to all the other train-mode outputs.  mixed: the ray bender runs on the tensor
    cores in the forward while the divergence pass recomputes it in fp32; alphas at the mode's own tolerance."""
    from gpu_common import build_composer
    config, _, inputs, comp, dev = build_composer(name, precision, training=True)
    comp.compute_divergence = True
    comp.allow_forward_without_grad = False
    e = [t.cuda() for t in scenes.divergence_noise(9, config, inputs)]
    res = comp(*[dev[k] for k in INPUT_KEYS], False, divergence_noise={"coarse": e})
    torch.cuda.synchronize()
    golden = load_golden(name + "_div")
    got = flatten(res)
    key = "coarse/global/integrated_divergence"
    assert float(np.abs(got[key]).max()) > 1e-3
    div_keys = [k for k in golden if k.endswith("integrated_divergence")]
    for k in div_keys:
        assert scale_rel_err(got[k], golden[k]) < tol, (k, scale_rel_err(got[k], golden[k]))
    if precision != "mixed":
        bad = compare(got, golden, tol)
        assert not bad, bad


def test_divergence_is_off_by_default_and_in_eval():
    from gpu_common import build_composer
    _, _, _, comp, dev = build_composer("toy_world", "fp32", training=True)
    comp.allow_forward_without_grad = False
    res = comp(*[dev[k] for k in INPUT_KEYS], False)
    assert float(res["coarse"]["global"]["integrated_divergence"].abs().max()) == 0.0
    comp.compute_divergence = True
    comp.eval()
    res = comp(*[dev[k] for k in INPUT_KEYS], False)
    assert float(res["coarse"]["global"]["integrated_divergence"].abs().max()) == 0.0


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_forward_expected_positions_backward_matches_reference_autograd(precision):
    """Gradients of <c1, expected_positions> + <c2, opacity> (seeded cotangents): ray-bender parameters and deformation code through
    the displacements, the field's parameters through the opacity, rays and pose through both (weights detached, :614).

    Pose gradient: the reference's forward_expected_positions scales the sample spacing by the OBJECT-space |R d| (it has re-bound
    ``ray_directions`` by then, :656/:687; forward() uses the world-space |d|), the kernels always by |d|.  For a rigid pose the values
    are identical and the two gradients differ by g R d d^T / |d| in the rotation block -- a component ORTHOGONAL to the rotation
    manifold (<R d d^T, R A> = d^T A d = 0 for antisymmetric A), which no rotation parametrisation can see.  The oracle restates the
    reference's form and is pinned on it (tests/test_oracle_golden.py); here the w2o gradient is compared on the tangent space: the
    translation column and the antisymmetric part of R^T dL/dR."""
    from gpu_common import build_composer
    from make_golden_expected import object_inputs
    from make_golden_div_cases import EXPECTED_GRAD_CASES, expected_loss
    golden = np.load(os.path.join(_HERE, "golden", "expected_positions_grad.npz"))
    names = ("ray_origins", "ray_directions", "focal_normals", "transformation_matrix_w2o", "style", "deformation", "object_in_scene")
    for name, k in EXPECTED_GRAD_CASES:
        _, _, inputs, comp, _ = build_composer(name, precision)
        comp.allow_forward_without_grad = False
        args = [t.cuda() for t in object_inputs(inputs, k)]
        leaves = {}
        for i, n in enumerate(names):
            if n in scenes.GRAD_INPUT_KEYS:
                args[i] = args[i].clone().requires_grad_(True)
                leaves[n] = args[i]
        exp, opacity = comp.forward_expected_positions(*args, k, False)["coarse"]
        loss = expected_loss(name, k, exp, opacity)
        loss.backward()
        torch.cuda.synchronize()
        ref_loss = float(golden[f"{name}/{k}/loss"])
        assert abs(float(loss.item()) - ref_loss) <= 3e-4 * max(1.0, abs(ref_loss)), (name, k)
        tol = 1e-3 if name == "toy_world" else 5e-2          # tennis_dense: gradient conditioning of the 10-octave fields (test_gpu_backward.py)
        params = dict(comp.named_parameters())
        for key in golden.files:
            if not key.startswith(f"{name}/{k}/") or key.endswith("/loss"):
                continue
            kind, n = key[len(f"{name}/{k}/"):].split("/", 1)
            if kind == "input":
                g = leaves[n].grad
                got = g.cpu().numpy() if g is not None else np.zeros(tuple(leaves[n].shape), np.float32)
            else:
                g = params[n].grad
                got = scenes.grad_subsample(n, g.cpu().numpy() if g is not None else np.zeros(tuple(params[n].shape), np.float32))
            ref = golden[key]
            if n == "transformation_matrix_w2o":
                R = args[3].detach().cpu().numpy().reshape(-1, 4, 4)[:, :3, :3]

                def tangent(g):
                    g = g.reshape(-1, 4, 4)
                    s = np.einsum("iba,ibc->iac", R, g[:, :3, :3])
                    return np.concatenate([(s - s.transpose(0, 2, 1)).reshape(-1), g[:, :3, 3].reshape(-1)])

                got, ref = tangent(got), tangent(ref)
            err = scale_rel_err(got, ref) if np.abs(ref).max() > 0 else float(np.abs(got).max())
            assert err <= tol, (name, k, key, err)
