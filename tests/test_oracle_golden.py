"""The CPU oracle (oracle/render_oracle.py) against outputs of the upstream reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import scenes
from helpers import INPUT_KEYS, compare, flatten, load_golden
from oracle import render_oracle as O

TOL = 2e-5   # fp32 CPU vs fp32 CPU, different summation order only

EVAL_SCENES = list(scenes.SCENES)


def _run(name, **kw):
    config, state, inputs = scenes.SCENES[name]()
    args = [inputs[k] for k in INPUT_KEYS]
    return config, state, inputs, O.composer_forward(config, state, *args, **kw)


@pytest.mark.parametrize("name", EVAL_SCENES)
def test_eval_matches_reference(name):
    _, _, _, res = _run(name, perturb=False)
    bad = compare(flatten(res), load_golden(name), TOL)
    assert not bad, bad


@pytest.mark.parametrize("name", ["cfg1", "tennis_dense", "minecraft_small"])
def test_perturb_matches_reference(name):
    config, state, inputs = scenes.SCENES[name]()
    rand, noise = scenes.perturbation_tensors(7, config, inputs)
    res = O.composer_forward(config, state, *[inputs[k] for k in INPUT_KEYS], perturb=True, rand=rand, noise=noise)
    bad = compare(flatten(res), load_golden(name + "_perturb"), 5e-5)
    assert not bad, bad


@pytest.mark.parametrize("name", ["cfg1", "static_small", "tennis_dense"])
def test_train_mode_batchnorm_matches_reference(name):
    new_stats = {}
    _, state, _, res = _run(name, perturb=False, training=True, new_stats=new_stats)
    golden = load_golden(name + "_train")
    # the Hutchinson divergence is random by construction (object_composer.py:597); documented deviation
    bad = compare(flatten(res), golden, 1e-4, skip=("integrated_divergence",))
    assert not bad, bad
    for k, ref in golden.items():
        if k.startswith("state/"):
            key = k[len("state/"):]
            got = new_stats.get(key, state[key]).numpy()
            np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-6, err_msg=key)


def test_known_answer_invariants():
    """chunked == unchunked in eval, opacity in [0,1], weights sum to opacity."""
    config, state, inputs = scenes.SCENES["cfg1"]()
    args = [inputs[k] for k in INPUT_KEYS]
    full = O.composer_forward(config, state, *args, perturb=False)
    chunked = O.batchified_composer_call(config, state, *args, perturb=False, samples_per_image_batching=100)
    g, c = full["coarse"]["global"], chunked["coarse"]["global"]
    torch.testing.assert_close(g["integrated_features"], c["integrated_features"], rtol=1e-5, atol=1e-6)  # BLAS blocking may differ per chunk size
    assert float(g["opacity"].min()) >= 0.0 and float(g["opacity"].max()) <= 1.0 + 1e-6
    torch.testing.assert_close(g["weights"].sum(-1), g["opacity"])


def test_decoder_feature_grids_layout():
    H, W, strides = 32, 64, [4, 8]
    R = sum((H // s) * (W // s) for s in strides)
    feats = torch.arange(R * 192, dtype=torch.float32).reshape(1, 1, 1, R, 192)
    g4, g8 = O.decoder_feature_grids(feats, strides, (H, W), [64, 128])
    assert g4.shape == (1, 1, 1, 64, 8, 16) and g8.shape == (1, 1, 1, 128, 4, 8)
    assert g4[0, 0, 0, 5, 1, 2] == feats[0, 0, 0, 1 * 16 + 2, 5]
    assert g8[0, 0, 0, 7, 3, 1] == feats[0, 0, 0, 8 * 16 + 3 * 8 + 1, 64 + 7]


GRAD_CASES = [("cfg1", False), ("static_small", False), ("tennis_dense", False), ("minecraft_small", False), ("toy_world", False),
              ("cfg1", True), ("tennis_dense", True), ("toy_world", True)]


@pytest.mark.parametrize("name,training", GRAD_CASES)
def test_gradients_match_reference_autograd(name, training):
    """Autograd through the oracle against gradients recorded from the upstream code's own graph (make_golden.py run_grad):
    every parameter and every differentiable input of ObjectComposer.forward."""
    from helpers import compare_grads
    golden = load_golden(f"{name}_grad_train" if training else f"{name}_grad")
    config, state, inputs = scenes.SCENES[name]()
    state = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v) for k, v in state.items()}
    inputs = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in inputs.items()}
    res = O.composer_forward(config, state, *[inputs[k] for k in INPUT_KEYS], perturb=False, training=training)["coarse"]
    loss = scenes.grad_loss(res, [str(k) for k in golden["loss_keys"]])
    assert abs(loss.item() - float(golden["loss"])) <= 2e-5 * max(1.0, abs(float(golden["loss"])))
    loss.backward()
    zeros = lambda t: np.zeros(tuple(t.shape), np.float32)
    got_in = {k: (inputs[k].grad.numpy() if inputs[k].grad is not None else zeros(inputs[k])) for k in scenes.GRAD_INPUT_KEYS}
    got_par = {k: (v.grad.numpy() if v.grad is not None else zeros(v)) for k, v in state.items() if v.is_floating_point() and "running_" not in k}
    bad = compare_grads(got_in, got_par, golden, 2e-4)
    assert not bad, bad


def test_oracle_expected_positions_match_the_reference():
    """forward_expected_positions (model/object_composer.py:624-722) of the oracle against outputs of the upstream method."""
    import os
    import numpy as np
    import torch
    from make_golden_expected import EXPECTED_CASES, object_inputs
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "expected_positions.npz"))
    for name, k in EXPECTED_CASES:
        config, state, inputs = scenes.SCENES[name]()
        with torch.no_grad():
            exp, opacity = O.forward_expected_positions(config, state, *object_inputs(inputs, k), k, False)["coarse"]
        for got, key in ((exp, "expected_positions"), (opacity, "opacity")):
            ref = golden[f"{name}/{k}/{key}"]
            assert float(np.abs(got.numpy() - ref).max()) <= 2e-5 * max(float(np.abs(ref).max()), 1e-6), (name, k, key)


# ---- hierarchical ("fine") pass: use_fine object models (object_composer.py:561-578, ray_helper.py:1320-1403) ----
@pytest.mark.parametrize("name", list(scenes.FINE_SCENES))
def test_fine_pass_matches_reference(name):
    config, state, inputs = scenes.FINE_SCENES[name]()
    res = O.composer_forward(config, state, *[inputs[k] for k in INPUT_KEYS], perturb=False)
    golden = load_golden(name)
    assert any(k.startswith("fine/") for k in golden)
    for family in ("coarse/", "fine/"):
        bad = compare(flatten(res), golden, 5e-5, only_prefix=family)
        assert not bad, bad


def test_fine_pass_train_mode_matches_reference():
    config, state, inputs = scenes.FINE_SCENES["toy_fine"]()
    new_stats = {}
    res = O.composer_forward(config, state, *[inputs[k] for k in INPUT_KEYS], perturb=False, training=True, new_stats=new_stats)
    golden = load_golden("toy_fine_train")
    for family in ("coarse/", "fine/"):
        bad = compare(flatten(res), golden, 1e-4, skip=("integrated_divergence",), only_prefix=family)
        assert not bad, bad
    fine_stats = [k for k in golden if k.startswith("state/object_models_fine.")]
    assert fine_stats
    for k, ref in golden.items():
        if k.startswith("state/"):
            key = k[len("state/"):]
            np.testing.assert_allclose(new_stats.get(key, state[key]).numpy(), ref, rtol=1e-4, atol=1e-6, err_msg=key)


@pytest.mark.parametrize("name", ["toy_world", "tennis_dense"])
def test_hutchinson_divergence_matches_reference(name):
    """compute_approximate_divergence (object_composer.py:582-601) with the probe vectors of the golden run
    (tests/golden/make_golden_div.py): integrated_divergence = mean(alpha |e . J e|) per object and for the composed scene."""
    config, state, inputs = scenes.SCENES[name]()
    e = scenes.divergence_noise(9, config, inputs)
    res = O.composer_forward(config, state, *[inputs[k] for k in INPUT_KEYS], perturb=False, training=True, new_stats={},
                             divergence_noise=e)
    golden = load_golden(name + "_div")
    assert float(np.abs(golden["coarse/global/integrated_divergence"]).max()) > 1e-3
    bad = compare(flatten(res), golden, 1e-4)
    assert not bad, bad


def test_oracle_expected_positions_gradients_match_the_reference():
    """Autograd through the oracle's forward_expected_positions against the upstream autograd (tests/golden/make_golden_div.py) --
    including the w2o gradient, whose rotation block carries the |R d| term of the object-space sample spacing (:656, :687)."""
    from make_golden_expected import object_inputs
    from make_golden_div_cases import expected_loss
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "expected_positions_grad.npz"))
    name, k = "toy_world", 2
    config, state, inputs = scenes.SCENES[name]()
    args = object_inputs(inputs, k)
    for i in (0, 1, 3, 4, 5):
        args[i] = args[i].clone().requires_grad_(True)
    exp, opacity = O.forward_expected_positions(config, state, *args, k, False)["coarse"]
    expected_loss(name, k, exp, opacity).backward()
    for i, n in ((0, "ray_origins"), (1, "ray_directions"), (3, "transformation_matrix_w2o"), (4, "style"), (5, "deformation")):
        ref = golden[f"{name}/{k}/input/{n}"]
        got = args[i].grad.numpy() if args[i].grad is not None else np.zeros_like(ref)
        assert np.abs(got - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-6), n
