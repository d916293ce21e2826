"""Helpers for the GPU parity tests: build the B200 composer for a seeded scene and run it."""
import copy

import torch

import scenes
from helpers import INPUT_KEYS
from playableenvironments_b200.model.object_composer import ObjectComposer


def build_composer(name_or_scene, precision="fp32", device="cuda", training=False):
    config, state, inputs = scenes.SCENES[name_or_scene]() if isinstance(name_or_scene, str) else name_or_scene
    comp = ObjectComposer(copy.deepcopy(config))
    missing, unexpected = comp.load_state_dict(state, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    comp.precision = precision
    comp = comp.to(device)
    comp.train(training)
    comp.allow_forward_without_grad = True
    dev_inputs = {k: v.to(device) for k, v in inputs.items()}
    return config, state, inputs, comp, dev_inputs


def run_composer(comp, dev_inputs, perturb=False, **kw):
    with torch.no_grad():
        return comp(*[dev_inputs[k] for k in INPUT_KEYS], perturb, **kw)
