"""Shared test helpers: golden loading and nested-dict comparison."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

INPUT_KEYS = ("ray_origins", "ray_directions", "focal_normals", "transformation_matrix_w2o",
              "style", "deformation", "object_in_scene")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def flatten(results, prefix=""):
    out = {}
    for k, v in results.items():
        if torch.is_tensor(v):
            out[prefix + k] = v.detach().cpu().numpy()
        elif isinstance(v, dict):
            out.update(flatten(v, prefix + k + "/"))
    return out


def scale_rel_err(got, ref):
    """max|got-ref| / max(|ref|max, tiny): error relative to the tensor's scale."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    # disparity is 0/0 = NaN on rays with zero opacity in the reference too: NaNs must coincide
    nan_ref, nan_got = np.isnan(ref), np.isnan(got)
    if not np.array_equal(nan_ref, nan_got):
        return float("inf")
    ref, got = ref[~nan_ref], got[~nan_ref]
    denom = max(float(np.abs(ref).max()) if ref.size else 0.0, 1e-12)
    return float(np.abs(got - ref).max() / denom) if ref.size else 0.0


def compare(flat_got, golden, tol, skip=(), only_prefix="coarse/"):
    """Returns {key: err} of entries above tol."""
    bad = {}
    for k, ref in golden.items():
        if not k.startswith(only_prefix) or any(s in k for s in skip):
            continue
        assert k in flat_got, f"missing output {k}"
        assert flat_got[k].shape == ref.shape, (k, flat_got[k].shape, ref.shape)
        e = scale_rel_err(flat_got[k], ref)
        if not (e <= tol):
            bad[k] = e
    return bad


def compare_grads(got_inputs, got_params, golden, tol):
    """got_inputs: {input name: array}, got_params: {state_dict name: array}; golden: a *_grad.npz dict
    (parameter gradients sub-sampled by scenes.grad_subsample).  Returns {key: err} above tol, errors relative to the
    golden tensor's max magnitude."""
    import scenes
    bad = {}
    for k, ref in golden.items():
        if k.startswith("input/"):
            got = got_inputs[k[len("input/"):]]
            assert got.shape == ref.shape, (k, got.shape, ref.shape)
        elif k.startswith("param/"):
            name = k[len("param/"):]
            got = scenes.grad_subsample(name, got_params[name])
        else:
            continue
        if not np.abs(ref).max() > 0:
            e = float(np.abs(got).max())
        else:
            e = scale_rel_err(got, ref)
        if not (e <= tol):
            bad[k] = e
    return bad


def free_port() -> int:
    """A TCP port that is free right now on 127.0.0.1 (rendezvous of the multi-process tests)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]
