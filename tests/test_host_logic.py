"""Host-side mirror of the reference interface: registry, state_dict names, helpers, ray sharding (CPU only)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenes
from helpers import free_port
from oracle import render_oracle as O
from playableenvironments_b200 import registry, sharding
from playableenvironments_b200.model.annealable_positional_encoder import annealing_weights
from playableenvironments_b200.model.object_composer import ObjectComposer
from playableenvironments_b200.utils.lib_3d.ray_helper import RayHelper
from playableenvironments_b200.utils.tensor_batchifier import TensorBatchifier
from playableenvironments_b200.utils.tensor_folder import TensorFolder


def test_registry_resolves_reference_architecture_strings():
    for name in ("model.nerf_models.ray_bending_style_nerf_model", "model.nerf_models.adain_style_nerf_model",
                 "model.nerf_models.skybox_adain_style_nerf_model_v3", "model.nerf_models.zeroed_ray_bender_model",
                 "model.nerf_models.positional_ray_bender_model"):
        module = registry.resolve(name)
        assert module.__name__ == "playableenvironments_b200." + name
        assert callable(getattr(module, "model"))


def test_state_dict_names_and_shapes_match_the_reference():
    """Checkpoints of the reference must load unchanged: same keys, same shapes."""
    for name in ("tennis_small", "minecraft_small", "cfg1"):
        config, state, _ = scenes.SCENES[name]()
        comp = ObjectComposer(copy.deepcopy(config))
        own = comp.state_dict()
        assert set(own.keys()) == set(state.keys())
        for k, v in state.items():
            assert tuple(own[k].shape) == tuple(v.shape), k
    shipped = ObjectComposer(copy.deepcopy(scenes.SCENES["tennis_small"]()[0]))
    assert sum(p.numel() for p in shipped.object_models_coarse[0].parameters()) == 666305     # SURVEY 8d cross-check
    assert sum(p.numel() for p in shipped.object_models_coarse[1].ray_bender.parameters()) == 101248


def test_composer_attributes_read_by_the_scene_model():
    config, state, _ = scenes.SCENES["minecraft_small"]()
    comp = ObjectComposer(copy.deepcopy(config))
    m = comp.object_models_coarse[2]
    assert m.model_config["positions_count_coarse"] == 32 and m.empty_space_alpha == -3.5
    assert m.bounding_box.get_size().tolist() == pytest.approx([1.2, 2.1, 2.4])
    h = comp.object_id_helper
    assert (h.objects_count, h.static_objects_count, h.dynamic_objects_count) == (4, 2, 2)
    assert [h.model_idx_by_object_idx(i) for i in range(4)] == [0, 1, 2, 2]
    assert comp.object_models_fine[0] is None


def test_wrong_object_count_raises_like_the_reference():
    config, state, inputs = scenes.SCENES["cfg1"]()
    comp = ObjectComposer(copy.deepcopy(config))
    bad = torch.eye(4).reshape(1, 1, 1, 4, 4, 1).repeat(1, 1, 1, 1, 1, 2)
    with pytest.raises(Exception, match="Transformation matrix must specifies"):
        comp(inputs["ray_origins"], inputs["ray_directions"], inputs["focal_normals"], bad, inputs["style"], inputs["deformation"],
             inputs["object_in_scene"], False)


def test_set_step_drives_the_annealing_weights():
    config, state, _ = scenes.SCENES["tennis_small"]()
    comp = ObjectComposer(copy.deepcopy(config))
    comp.load_state_dict(state)
    enc = comp.object_models_coarse[1].ray_bender.positional_encoder
    assert enc.host_step() == 60000
    comp.set_step(21000)
    assert enc.host_step() == 21000 and int(enc.current_step) == 21000
    torch.testing.assert_close(annealing_weights(21000, 6, 60000), O.annealing_weights(21000, 6, 60000))


def test_tensor_helpers():
    x = torch.arange(2 * 3 * 7 * 5).reshape(2, 3, 7, 5)
    chunks = TensorBatchifier.batchify(x, dim=-2, batch_size=3)
    assert [c.size(-2) for c in chunks] == [3, 3, 1] and torch.equal(torch.cat(chunks, dim=-2), x)
    flat, dims = TensorFolder.flatten(x, -1)
    assert flat.shape == (42, 5) and dims == [2, 3, 7]
    assert torch.equal(TensorFolder.fold(flat, dims), x)
    with pytest.raises(Exception):
        TensorFolder.fold(flat, [5])


def test_ray_helper_shape_functions_match_the_oracle():
    H, W = 16, 24
    focal = torch.tensor([[30.0, 41.0]])
    d, o, n = RayHelper.create_camera_rays([1, 2], H, W, focal)
    rd, ro, rn = O.create_camera_rays([1, 2], H, W, focal)
    assert torch.equal(d, rd) and torch.equal(o, ro) and torch.equal(n, rn)
    obs = torch.rand(1, 2, 3, H, W)
    sd, so, sp = RayHelper.sample_all_rays_strided_grid(d, obs, [4, 8])
    od, op = O.sample_all_rays_strided_grid(rd, [4, 8])
    assert torch.equal(sd, od) and torch.allclose(sp, op)
    assert so.shape == (1, 2, sd.size(-2), 3)
    folded = RayHelper.fold_strided_grid_samples(sd, [4, 8], (H, W), dim=-2)
    assert [tuple(f.shape) for f in folded] == [(1, 2, 4, 6, 3), (1, 2, 2, 3, 3)]
    with pytest.raises(Exception, match="not divisible"):
        RayHelper.sample_strided_grid(d, 5)
    c2w = torch.from_numpy(scenes.tennis_camera()).float().expand(1, 2, 4, 4)
    to, td, tn = RayHelper.transform_rays(o, sd, n, c2w)
    oo, od2, on = O.transform_rays(ro, od, rn, c2w)
    torch.testing.assert_close(td, od2)
    torch.testing.assert_close(to, oo)


def test_ray_shards_partition_the_rays():
    for rays, world, mult in [(65536, 8, 1), (11520, 8, 4), (10, 4, 1), (3, 8, 1), (1000, 3, 128)]:
        spans = [sharding.ray_shard(rays, r, world, mult) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == rays
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert sum(sharding.shard_sizes(rays, world, mult)) == rays


def _gather_worker(rank, world, port, rays, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(2 * rays * 5, dtype=torch.float32).reshape(2, rays, 5)
    b, e = sharding.ray_shard(rays, rank, world)
    out = sharding.all_gather_rays(full[:, b:e].contiguous(), rays, dim=-2)
    ok = torch.equal(out, full)
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("rays", [11, 64])
def test_all_gather_of_ray_shards_world_size_2(tmp_path, rays):
    """The single collective of the path (feature-grid all-gather) on the gloo backend, 2 ranks, uneven shards."""
    port = free_port()
    mp.spawn(_gather_worker, args=(2, port, rays, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]


def _allreduce_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    params = [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.randn(3, 3)),
              torch.nn.Parameter(torch.randn(2), requires_grad=False)]
    grads = [[torch.full_like(p, float(r + 1)) * (i + 1) for i, p in enumerate(params)] for r in range(world)]
    params[0].grad, params[1].grad = grads[rank][0].clone(), grads[rank][1].clone()      # params[2]: no gradient on any rank but rank 1
    if rank == 1:
        params[2].grad = grads[1][2].clone()
    nbytes = sharding.allreduce_gradients(params, average=True)
    want0 = sum(grads[r][0] for r in range(world)) / world
    want1 = sum(grads[r][1] for r in range(world)) / world
    want2 = grads[1][2] / world
    ok = (nbytes == (35 + 11 + 9) * 4 and torch.allclose(params[0].grad, want0) and torch.allclose(params[1].grad, want1)
          and torch.allclose(params[2].grad, want2) and params[3].grad is None)
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2(tmp_path):
    """Data-parallel training step: one flat bucket, one all-reduce, averaged like nn.DataParallel's replica reduction (train.py:61)."""
    port = free_port()
    mp.spawn(_allreduce_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]


def test_fine_ray_parameters_match_the_oracle():
    """Host side of the fine pass (ObjectComposer._fine_ray_parameters: torch z-bounds + coarse ray parameters + inverse-CDF resampling
    + merge) against the oracle's restatement of create_ray_positions_weighted, itself pinned to the upstream goldens."""
    import copy
    import scenes
    from helpers import INPUT_KEYS
    from oracle import render_oracle as O
    from playableenvironments_b200.model.object_composer import ObjectComposer
    config, state, inputs = scenes.FINE_SCENES["toy_fine"]()
    comp = ObjectComposer(copy.deepcopy(config))
    assert comp._uses_fine()
    args = [inputs[k] for k in INPUT_KEYS]
    ref = O.composer_forward(config, state, *args, perturb=False)
    ts = comp._fine_ray_parameters(inputs["ray_origins"], inputs["ray_directions"], inputs["focal_normals"],
                                   inputs["transformation_matrix_w2o"], inputs["object_in_scene"], False, None, ref["coarse"])
    model_of, _ = O.object_ids(config)
    for k, t in enumerate(ts):
        cfg = config["model"]["object_models"][model_of[k]]
        assert t.shape[-1] == cfg["positions_count_coarse"] + cfg["positions_count_fine"]
        assert bool((t[..., 1:] >= t[..., :-1]).all())
        # the fine object's weights live on exactly these ray parameters: depth = sum w t
        w = ref["fine"][f"object_{k}"]["weights"]
        torch.testing.assert_close((w * t).sum(-1), ref["fine"][f"object_{k}"]["depth"], rtol=1e-4, atol=1e-5)


def test_reference_arm_zip_is_the_upstream_composer():
    """bench.py's reference arms (cpu_baseline / --impl reference / gpu_eager_baseline) run the UPSTREAM composer out of
    oracle/_ref/reference_path.zip (built by oracle/make_ref.py in the build container).  In a subprocess (its CPU shims are global):
    the zip imports and reproduces a committed golden (2e-5: BLAS threading may reorder sums)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "reference_path.zip")):
        pytest.skip("oracle/_ref/reference_path.zip not built (python oracle/make_ref.py in the build container)")
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path[:0] = [{root!r}, {os.path.join(root, 'tests')!r}, {os.path.join(root, 'tests', 'golden')!r}]\n"
        "import bench, scenes\n"
        "from helpers import INPUT_KEYS, load_golden\n"
        "config, state, inputs = scenes.SCENES['cfg1']()\n"
        "comp = bench.upstream_composer(config, state).eval()\n"
        "assert type(comp).__module__ == 'model.object_composer' and 'reference_path.zip' in sys.modules['model.object_composer'].__file__\n"
        "with torch.no_grad():\n"
        "    out = comp(*[inputs[k] for k in INPUT_KEYS], False)['coarse']['global']['integrated_features'].numpy()\n"
        "ref = load_golden('cfg1')['coarse/global/integrated_features']\n"
        "assert np.abs(out - ref).max() <= 2e-5 * np.abs(ref).max(), np.abs(out - ref).max()\n"
        "print('ok')\n")
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and "ok" in proc.stdout, proc.stderr[-800:]


def test_hand_off_plan_and_tile_count_policy(monkeypatch):
    """Host-side decisions of the render call: which decoder hand-off plans the compositor can write directly (rays = the concatenated
    strided grids, channel ranges in multiples of 32), and when the forward counts the backward's tiles (only when the worst case
    overflows the activation stash)."""
    from playableenvironments_b200.model import render
    plan = ([4, 8], (288, 512), [64, 128])
    assert render.handoff_supported(plan, 72 * 128 + 36 * 64, 192)
    assert not render.handoff_supported(plan, 72 * 128, 192)                         # rays are not the two grids
    assert not render.handoff_supported(([4, 8], (288, 512), [64, 100]), 11520, 192)   # channel range not a multiple of 32
    assert not render.handoff_supported(([4, 8], (288, 512), [128, 128]), 11520, 192)  # more channels than features
    assert not render.handoff_supported(([4, 8, 16, 32, 64], (288, 512), [32] * 5), 0, 192)

    class D:
        def __init__(self, positions):
            self.positions = positions
    monkeypatch.delenv("PE_BWD_TILE_COUNTS", raising=False)
    monkeypatch.delenv("PE_BWD_TC_MAX_TILES", raising=False)
    tennis = [D(4), D(4), D(32), D(32)]
    assert not render._wants_tile_counts(tennis, 4, 5120)          # one replica: 5 120 tiles per player at most
    assert render._wants_tile_counts(tennis, 32, 5120)             # the 32-image batch: 40 960 > 12 288
    monkeypatch.setenv("PE_BWD_TC_MAX_TILES", "100000")
    assert not render._wants_tile_counts(tennis, 32, 5120)
    monkeypatch.setenv("PE_BWD_TILE_COUNTS", "1")
    assert render._wants_tile_counts(tennis, 1, 8)
    monkeypatch.setenv("PE_BWD_TILE_COUNTS", "0")
    assert not render._wants_tile_counts(tennis, 32, 5120)
