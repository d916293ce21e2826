"""Ray selection: the tensorised, sync-free mirror of the reference's RayHelper.sample_rays* against golden outputs
of the upstream functions (tests/golden/make_golden_rays.py) for the same torch RNG seed.  Index work: bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(HERE, "golden")]
from make_golden_rays import fine_case, ray_cases  # noqa: E402
from playableenvironments_b200.utils.lib_3d.ray_helper import RayHelper  # noqa: E402

GOLDEN = np.load(os.path.join(HERE, "golden", "ray_selection.npz"))
CASES = ray_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_the_reference_bit_exactly(name):
    fn, seed, kwargs = CASES[name]
    torch.manual_seed(seed)
    res = getattr(RayHelper, fn)(**kwargs)
    for i, t in enumerate(res):
        want = GOLDEN[f"{name}/{i}"]
        assert tuple(t.shape) == want.shape, (name, i, t.shape, want.shape)
        np.testing.assert_array_equal(t.numpy(), want, err_msg=f"{name} output {i}")


def test_strided_patch_rays_are_cell_centres_inside_the_image():
    fn, seed, kwargs = CASES["strided_patch_tennis_like"]
    torch.manual_seed(99)
    _, _, pos = RayHelper.sample_rays_strided_patch(**kwargs)
    h, w = kwargs["ray_directions"].shape[-3:-1]
    splits = RayHelper.split_strided_patch_ray_samples(pos, kwargs["patch_size"], kwargs["strides"])
    for stride, p in zip(kwargs["strides"], splits):
        rows = (p[..., 0] * h).round().long()
        cols = (p[..., 1] * w).round().long()
        assert ((rows % stride) == stride // 2).all() and ((cols % stride) == stride // 2).all()
        assert (rows >= 0).all() and (rows < h).all() and (cols >= 0).all() and (cols < w).all()


def test_explicit_uniform_numbers_replace_the_generator():
    fn, seed, kwargs = CASES["weighted"]
    u = torch.rand((4, 200), generator=torch.Generator().manual_seed(5))
    a = RayHelper.sample_rays_weighted(**kwargs, cdf_samples=u)
    b = RayHelper.sample_rays_weighted(**kwargs, cdf_samples=u)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_argument_errors_match_the_reference():
    fn, seed, kwargs = CASES["strided_patch_2strides"]
    with pytest.raises(Exception, match="Align grid"):
        RayHelper.sample_rays_strided_patch(**{**kwargs, "align_grid": False})
    with pytest.raises(Exception, match="multiple of 2"):
        RayHelper.sample_rays_strided_patch(**{**kwargs, "patch_size": 7})
    with pytest.raises(Exception, match="not compatible"):
        RayHelper.sample_rays_strided_patch(**{**kwargs, "patch_size": 2})


def test_install_ray_selection_rebinds_a_reference_style_class():
    from playableenvironments_b200.model.environment_model_glue import install_ray_selection, RAY_SELECTION_FUNCTIONS

    class FakeReferenceRayHelper:          # stands in for the upstream class (not importable on the GPU box)
        pass

    install_ray_selection(FakeReferenceRayHelper)
    for name in RAY_SELECTION_FUNCTIONS:
        assert getattr(FakeReferenceRayHelper, name) is getattr(RayHelper, name)


@pytest.mark.parametrize("perturb", [False, True])
def test_weighted_fine_sampling_matches_the_reference(perturb):
    """create_ray_positions_weighted / sample_pdf (reference :1320-1403) for the same RNG state."""
    origins, dirs, t, w = fine_case()
    torch.manual_seed(33)
    pos, merged = RayHelper.create_ray_positions_weighted(origins, dirs, 24, t, w, perturb)
    np.testing.assert_allclose(merged.numpy(), GOLDEN[f"fine_{int(perturb)}/1"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(pos.numpy(), GOLDEN[f"fine_{int(perturb)}/0"], rtol=0, atol=2e-5)
    assert torch.equal(w, fine_case()[3])          # the argument is left untouched


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_the_reference_on_the_device(name, monkeypatch):
    """The same selection with every tensor on the GPU.  torch's CUDA generator is a different stream than the CPU one the goldens were
    drawn from, so the random numbers are drawn on the host exactly as above and moved: everything downstream (cdf inversion,
    searchsorted, index arithmetic, gathers) runs on the device and must reproduce the upstream results: bit for bit where the output is
    a gather of the inputs or an index, to 1 ulp for the normalised pixel positions (ATen's CUDA kernel divides by a scalar as a
    multiplication with its reciprocal)."""
    fn, seed, kwargs = CASES[name]
    rand, randperm = torch.rand, torch.randperm

    def host_rand(*size, device=None, **kw):
        return rand(*size, **kw).to(device) if device is not None else rand(*size, **kw)

    def host_randperm(n, device=None, **kw):
        return randperm(n, **kw).to(device) if device is not None else randperm(n, **kw)

    monkeypatch.setattr(torch, "rand", host_rand)
    monkeypatch.setattr(torch, "randperm", host_randperm)
    dev_kwargs = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kwargs.items()}
    torch.manual_seed(seed)
    res = getattr(RayHelper, fn)(**dev_kwargs)
    for i, t in enumerate(res):
        assert t.is_cuda, (name, i)
        got, want = t.cpu().numpy(), GOLDEN[f"{name}/{i}"]
        if got.dtype.kind == "f" and got.shape[-1] == 2 and float(np.abs(want).max()) <= 1.0:      # normalised (row, column) positions
            np.testing.assert_allclose(got, want, rtol=2.5e-7, atol=0, err_msg=f"{name} output {i}")
        else:
            np.testing.assert_array_equal(got, want, err_msg=f"{name} output {i}")
