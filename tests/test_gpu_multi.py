"""Multi-GPU product path on NCCL (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
`sharding.render_sharded` / `render_pipelined` against the single-GPU render of the same frame, bit for bit (rays are independent),
and the data-parallel gradient all-reduce (one flat bucket) against a per-tensor reduction."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE, os.path.join(_HERE, "golden")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from helpers import free_port  # noqa: E402

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    import scenes
    from helpers import INPUT_KEYS
    from gpu_common import build_composer
    from playableenvironments_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    ok = []
    try:
        for scene, precision in ((scenes.scene_static(seed=12, height=40, width=50, P=128), "mixed"),
                                 (scenes.scene_tennis(seed=13, stride=8, lead=(1, 1, 1)), "fp16x3")):
            _, _, _, comp, dev = build_composer(scene, precision, device=device)
            args = [dev[k] for k in INPUT_KEYS]
            rays = dev["ray_directions"].size(-2)
            with torch.no_grad():
                single = comp(*args, False)["coarse"]["global"]["integrated_features"]
                full, local, (b, e) = sharding.render_sharded(comp, *args, False)
                ok.append(torch.equal(full, single))
                ok.append(torch.equal(local["coarse"]["global"]["integrated_features"], single[..., b:e, :]))
                # weak-scaling form: every rank renders the whole frame in 3 pipelined chunks, grids gathered on a side stream
                F = single.size(-1)
                gathered = torch.zeros((world, rays, F), device=device)
                host = torch.zeros((rays, F)).pin_memory()
                mine = sharding.render_pipelined(comp, *args, False, chunks=3, gathered=gathered, host_out=host)
                torch.cuda.synchronize()
                ok.append(torch.equal(mine, single.reshape(rays, F)))
                ok.append(all(torch.equal(gathered[r], single.reshape(rays, F)) for r in range(world)))
                ok.append(torch.equal(host, single.reshape(rays, F).cpu()))
                # the all-gather fused into the render kernel (P2P stores into every rank's grid), single launch and pipelined with D2H
                pg = sharding.PeerGather(rays, F, device)
                comp(*args, False, peer_features=pg.destinations())
                grid = pg.sync()
                torch.cuda.synchronize()
                ok.append(all(torch.equal(grid[r], single.reshape(rays, F)) for r in range(world)))
                pg.buffer.zero_()
                dist.barrier()
                host.zero_()
                sharding.render_pipelined(comp, *args, False, chunks=3, host_out=host, peer_gather=pg)
                torch.cuda.synchronize()
                ok.append(all(torch.equal(pg.buffer[r], single.reshape(rays, F)) for r in range(world)))
                ok.append(torch.equal(host, single.reshape(rays, F).cpu()))
        # data-parallel training: rank r differentiates its own frame; the single flat-bucket all-reduce must equal a per-tensor reduction
        _, _, _, mine_comp, dev = build_composer(scenes.scene_static(seed=40 + rank, height=8, width=8, P=64), "fp32", device=device)
        mine_comp.allow_forward_without_grad = False
        mine_comp(*[dev[k] for k in INPUT_KEYS], False)["coarse"]["global"]["integrated_features"].sum().backward()
        before = {k: p.grad.clone() for k, p in mine_comp.named_parameters() if p.grad is not None}
        nbytes = sharding.allreduce_gradients(mine_comp.parameters(), average=False)
        gathered_sum = {}
        for k, g in before.items():
            t = g.clone()
            dist.all_reduce(t)
            gathered_sum[k] = t
        ok.append(nbytes > 0 and all(torch.allclose(p.grad, gathered_sum[k], rtol=1e-6, atol=0) for k, p in mine_comp.named_parameters() if k in gathered_sum))
    finally:
        open(os.path.join(tmp, f"ok{rank}"), "w").write(",".join("1" if x else "0" for x in ok))
        dist.destroy_process_group()


def test_sharded_and_pipelined_render_on_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        flags = open(tmp_path / f"ok{r}").read().split(",")
        assert flags and all(f == "1" for f in flags), (r, flags)


def test_data_parallel_replicas_on_two_gpus():
    """The reference's own multi-GPU wrapper (train.py:61: nn.DataParallel): the batch dimension scattered over two replicas running in two
    Python threads, one device each -- every object of a replica on its own stream of that device's pool.  Eval-mode images are
    independent, so outputs equal the single-device call image by image and the reduced parameter gradients equal its gradients."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import scenes
    from helpers import INPUT_KEYS
    from gpu_common import build_composer
    scene = scenes.scene_tennis(seed=13, stride=8, lead=(2, 2, 1))
    _, _, _, comp, dev = build_composer(scene, "mixed", device=torch.device("cuda", 0))
    comp.allow_forward_without_grad = False
    args = [dev[k] for k in INPUT_KEYS]

    def loss_of(res):
        return (res["coarse"]["global"]["integrated_features"].sum() + res["coarse"]["global"]["opacity"].sum()
                + res["coarse"]["object_1"]["depth"].sum())

    single = comp(*args, False)
    loss_of(single).backward()
    ref_grads = {k: p.grad.clone() for k, p in comp.named_parameters() if p.grad is not None}
    comp.zero_grad(set_to_none=True)
    parallel = torch.nn.DataParallel(comp, device_ids=[0, 1])
    for _ in range(2):                                   # (second pass: cached packed weights and stream pools of both devices)
        comp.zero_grad(set_to_none=True)
        both = parallel(*args, False)
        loss_of(both).backward()
    torch.cuda.synchronize()
    for name in ("global", "object_0", "object_1", "object_2"):
        for key in ("integrated_features", "opacity", "depth", "weights"):
            assert torch.equal(both["coarse"][name][key], single["coarse"][name][key]), (name, key)
    assert ref_grads
    for k, p in comp.named_parameters():
        if k in ref_grads:
            scale = float(ref_grads[k].abs().max())
            assert float((p.grad - ref_grads[k]).abs().max()) <= 2e-4 * max(scale, 1e-12), k
