"""How well-conditioned are the GOLDEN gradients themselves?  The oracle evaluates the same graph in float64; the distance of the
reference's own fp32 autograd gradients (tests/golden/*_grad.npz) from that evaluation is the floor any fp32 implementation can be
held to -- it justifies the per-scene tolerances of tests/test_gpu_backward.py (static_small 8e-3, minecraft_small 1e-2,
tennis_dense 5e-2: d sin(512 x)/dx multiplies rounding noise by 512 and sums over ~1e5 samples cancel heavily)."""
import numpy as np
import pytest
import torch

import scenes
from helpers import INPUT_KEYS, compare_grads, load_golden
from oracle import render_oracle as O


def float64_gradient_distance(name):
    golden = load_golden(f"{name}_grad")
    config, state, inputs = scenes.SCENES[name]()
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)       # the oracle's factory calls (linspace, zeros, ...) follow the default dtype
    try:
        state64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in state.items()}
        args = {k: (v.double() if v.is_floating_point() else v) for k, v in inputs.items()}
        for k in scenes.GRAD_INPUT_KEYS:
            args[k] = args[k].clone().requires_grad_(True)
        res = O.composer_forward(config, state64, *[args[k] for k in INPUT_KEYS], perturb=False)["coarse"]
        scenes.grad_loss(res, [str(k) for k in golden["loss_keys"]]).backward()
    finally:
        torch.set_default_dtype(old)
    got_in = {k: (args[k].grad.float().numpy() if args[k].grad is not None else np.zeros(tuple(args[k].shape), np.float32)) for k in scenes.GRAD_INPUT_KEYS}
    got_par = {k: (v.grad.float().numpy() if (torch.is_tensor(v) and v.grad is not None) else np.zeros(tuple(v.shape), np.float32))
               for k, v in state64.items() if "running_" not in k and v.is_floating_point()}
    errs = compare_grads(got_in, got_par, golden, -1.0)
    return max(errs.values()), max(errs, key=errs.get)


@pytest.mark.parametrize("name,low,high", [("static_small", 5e-4, 8e-3), ("minecraft_small", 5e-4, 1e-2), ("tennis_dense", 5e-2, 1.0)])
def test_reference_fp32_gradients_against_float64(name, low, high):
    worst, key = float64_gradient_distance(name)
    assert low <= worst <= high, (name, worst, key)
