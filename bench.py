"""Benchmark of the per-frame NeRF render path (BASELINE.json metric: ray-samples/sec of the fused style-MLP ray-march
at 256x256 rays x 128 samples/ray).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision mixed|fp16|fp16x2|fp16x3|fp32] [--impl b200|reference]

A step = one pass of the hot path over one synthetic frame per GPU (BASELINE configs[1]: one static field, shipped
8x256/192-feature architecture, camera inside the box so all 8 388 608 samples are in-box), followed for N > 1 by the
single all-gather of the rendered feature grids (weak scaling: one frame per rank).  Prints ONE JSON line (rank 0).

`--impl reference` times the reference's CPU implementation of the same path — the oracle port of oracle/render_oracle.py
(the upstream code is Python/PyTorch and is not shipped to the GPU box; the port is pinned against it by
tests/golden) — on the host cores, on a bounded ray subset of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]

FLOP_PER_SAMPLE = 2 * 614144          # matmul MACs of the shipped field x 2
HEIGHT, WIDTH, POSITIONS = 256, 256, 128
TRAFFIC_MIXED = 65513984          # dram bytes of one pe_field_tc_kernel launch in mixed mode (profiles/r2_aware_rounding.md: 2.82 MB read + 62.69 MB written)
WORKLOAD = "cfg2: 1 static field (W=256,L=8,skip=4,10 oct,F=192), 256x256 rays x 128 samples/ray, 100% in-box, forward"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"burst": p["bf16_tflops"], "sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def build_scene():
    import scenes
    return scenes.scene_static(seed=12, height=HEIGHT, width=WIDTH, P=POSITIONS)


UPSTREAM_ZIP = os.path.join(ROOT, "oracle", "_ref", "reference_path.zip")


def upstream_composer_class(cpu: bool):
    """The UPSTREAM ObjectComposer out of oracle/_ref/reference_path.zip (import closure of the reference's render path, built by
    oracle/make_ref.py in the build container; git-ignored, travels with the snapshot) or None when the zip is absent."""
    if not os.path.exists(UPSTREAM_ZIP):
        return None
    from oracle import make_ref
    make_ref.install_shims(cpu)
    if UPSTREAM_ZIP not in sys.path:
        sys.path.insert(0, UPSTREAM_ZIP)
    from model.object_composer import ObjectComposer
    return ObjectComposer


def upstream_composer(config, state, device=None):
    import copy
    cls = upstream_composer_class(cpu=device is None)
    if cls is None:
        return None
    comp = cls(copy.deepcopy(config))
    missing, unexpected = comp.load_state_dict(state, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return comp if device is None else comp.to(device)


def chunked_call(comp, args, chunk: int):
    """The caller's chunking (EnvironmentModel.batchified_composer_call, environment_model.py:474-521): rays in chunks of `chunk`."""
    rays = args[1].size(-2)
    out = []
    for b in range(0, rays, chunk):
        a = list(args)
        a[1] = args[1][..., b:b + chunk, :]
        out.append(comp(*a, False)["coarse"]["global"]["integrated_features"])
    return out


def cpu_reference(sample_rays: int, reps: int, warmup: int):
    """Times the reference path on the CPU on `sample_rays` rays of the workload, all host threads: the upstream ObjectComposer itself
    when oracle/_ref holds it (kind "reference"), else the oracle port (kind "port")."""
    import scenes  # noqa: F401
    from helpers import INPUT_KEYS
    from oracle import render_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    config, state, inputs = build_scene()
    inputs = dict(inputs)
    inputs["ray_directions"] = inputs["ray_directions"][..., :sample_rays, :].contiguous()
    args = [inputs[k] for k in INPUT_KEYS]
    comp = upstream_composer(config, state)
    if comp is not None:
        comp.eval()
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            # the reference bounds activation memory with samples_per_image_batching=1000 (environment_model.py:584)
            if comp is not None:
                chunked_call(comp, args, 1000)
            else:
                O.batchified_composer_call(config, state, *args, perturb=False, samples_per_image_batching=1000)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    samples = sample_rays * POSITIONS
    return samples, times, ("reference" if comp is not None else "port")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_rays = 2048
    samples, times, kind = cpu_reference(sample_rays, max(args.steps, 1), min(args.warmup, 1))
    mean = sum(times) / len(times)
    value = samples / mean
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "ray-samples/sec (style-MLP ray-march, fwd)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample_rays} rays x {POSITIONS} samples of the frame per step"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{sample_rays} of 65536 rays ({samples} samples) per step, chunks of 1000 rays like the reference"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def gpu_eager_reference(device, tennis_hw=(144, 256)):
    """SURVEY 8(d) / BASELINE.md section 3: the reference's own eager PyTorch formulation of the path (the oracle port: the same
    torch op sequence as upstream, pinned against it by tests/golden) timed ON THIS GPU, in plain fp32 and with TF32 matmuls:
    cfg2 forward in chunks of 8192 rays (the reference bounds activation memory by chunking, environment_model.py:584) and the cfg3
    Tennis train step (forward + backward through autograd, train-mode BatchNorm).  What the repo's kernels must beat on equal hardware."""
    import scenes
    from helpers import INPUT_KEYS
    from oracle import render_oracle as O
    real = upstream_composer_class(cpu=False) is not None
    out = {"kind": ("reference (the upstream ObjectComposer itself, oracle/_ref/reference_path.zip, eager on this GPU; its train step includes "
                    "its Hutchinson divergence pass, object_composer.py:582-601)") if real else
           "port (oracle/render_oracle.py under torch.device(cuda): ATen / cuBLAS kernels, like the upstream eager path)"}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            fn()
        e0.record()
        torch.cuda.synchronize()
        return s0.elapsed_time(e0) / reps

    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        with torch.device(device):
            config, state, inputs = build_scene()
            state_d = {k: v.to(device) for k, v in state.items()}
            args = [inputs[k].to(device) for k in INPUT_KEYS]

            comp2 = upstream_composer(config, state, device) if real else None
            if comp2 is not None:
                comp2.eval()

            def fwd():
                with torch.no_grad():
                    if comp2 is not None:
                        chunked_call(comp2, args, 8192)
                    else:
                        O.batchified_composer_call(config, state_d, *args, perturb=False, samples_per_image_batching=8192)

            tscene = scenes.scene_tennis(seed=13, height=tennis_hw[0], width=tennis_hw[1], stride=1, lead=(1, 1, 1), dense=True)
            tconfig, tstate, tinputs = tscene
            tstate_d = {k: v.to(device).requires_grad_(v.is_floating_point() and "running_" not in k) for k, v in tstate.items()}
            targs = [tinputs[k].to(device) for k in INPUT_KEYS]
            targs = [a.requires_grad_(True) if (k in scenes.GRAD_INPUT_KEYS) else a for k, a in zip(INPUT_KEYS, targs)]
            rays = tinputs["ray_directions"].size(-2)
            cot = torch.randn(rays, 192)

            comp3 = upstream_composer(tconfig, tstate, device) if real else None
            if comp3 is not None:
                comp3.train()

            def train():
                for t in list(tstate_d.values()) + targs:
                    t.grad = None
                if comp3 is not None:
                    comp3.zero_grad(set_to_none=True)
                    res = comp3(*targs, False)["coarse"]["global"]
                else:
                    res = O.composer_forward(tconfig, tstate_d, *targs, perturb=False, training=True)["coarse"]["global"]
                loss = (res["integrated_features"].reshape(rays, 192) * cot).sum() + res["opacity"].sum()
                loss.backward()

            for tf32 in (False, True):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                tag = "tf32" if tf32 else "fp32"
                ms = timed(fwd, 2)
                out[f"cfg2_fwd_{tag}_ms"] = ms
                out[f"cfg2_fwd_{tag}_samples_per_s"] = HEIGHT * WIDTH * POSITIONS / (ms / 1e3)
                out[f"cfg3_dense_train_step_{tag}_ms"] = timed(train, 2)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


def train_step_report(device, dense: bool, precision: str = "mixed", world: int = 1):
    """BASELINE configs[2]: Tennis scene (court + 2 players with positional ray benders, composed), 256x144 rays, forward and
    forward+backward through ObjectComposer with every parameter and every differentiable input requiring a gradient.
    Secondary figures (the headline stays configs[1]).  precision mixed (the composer's default): train-mode forward on the tensor
    cores, kept for the backward; field and ray-bender backward on the tensor cores (recompute with stash -> dX chain -> dW);
    precision fp32: everything on the CUDA cores (round 1's path).  world > 1: one frame per rank + ONE all-reduce of the flat
    gradient bucket (data-parallel step, train.py:61)."""
    import scenes
    from helpers import INPUT_KEYS
    from gpu_common import build_composer
    scene = scenes.scene_tennis(seed=13, height=144, width=256, stride=1, lead=(1, 1, 1), dense=dense)
    if precision == "fp32":
        os.environ["PE_TC_BACKWARD_RECOMPUTE"] = "0"
        os.environ["PE_BWD_TC"] = "0"
    config, state, inputs, comp, dev = build_composer(scene, precision, device=device, training=True)
    comp.allow_forward_without_grad = False
    dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
    call = [dev[k] for k in INPUT_KEYS]
    rays = dev["ray_directions"].size(-2)
    cot = torch.randn(rays, 192, device=device)

    def fwd():
        with torch.no_grad():
            return comp(*call, False)

    from playableenvironments_b200 import sharding
    params = list(comp.parameters())

    def fwd_bwd():
        comp.zero_grad(set_to_none=True)
        res = comp(*call, False)["coarse"]
        loss = (res["global"]["integrated_features"].reshape(rays, 192) * cot).sum() + res["global"]["opacity"].sum()
        loss.backward()
        if world > 1:           # data-parallel step: one frame per rank, ONE all-reduce of the flat gradient bucket (train.py:61)
            sharding.allreduce_gradients(params, average=True)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            fn()
        e0.record()
        torch.cuda.synchronize()
        return s0.elapsed_time(e0) / reps

    comp.return_raw_alphas = True
    res = fwd()["coarse"]
    comp.return_raw_alphas = False
    slots, inbox = 0, 0
    for k, m in enumerate(config["model"]["object_models"]):
        ra = res[f"object_{k}"]["raw_alphas"]
        slots += ra.numel()
        inbox += int((ra != float(m["empty_space_alpha"])).sum().item())
    ms_f, ms_fb = timed(fwd, 3), timed(fwd_bwd, 3)
    os.environ.pop("PE_TC_BACKWARD_RECOMPUTE", None)
    os.environ.pop("PE_BWD_TC", None)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_f, ms_fb], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_f, ms_fb = float(t[0]), float(t[1])
        inbox *= world
    return {"n_gpus": world, "gradient_allreduce_bytes": sum(p.numel() for p in params) * 4 if world > 1 else 0,"workload": f"cfg3 Tennis{' (dense: camera on a player)' if dense else ''}: court P=4 + 2 players P=32 with ray benders, 256x144 rays, train-mode BatchNorm",
            "sample_slots": slots, "in_box_samples": inbox, "fwd_ms": ms_f, "fwd_bwd_ms": ms_fb,
            "in_box_samples_per_s_fwd_bwd": inbox / (ms_fb / 1e3),
            "precision": "fp32 (CUDA cores only; the backward recomputes the forward)" if precision == "fp32"
            else f"{precision} forward, its workspace kept for the backward; field and ray-bender backward on the tensor cores (pe_bwd_tc.cu)"}


def eval_frame_report(device, dense: bool):
    """BASELINE configs[2] scene in EVAL mode (what play.py renders): court + 2 players with ray benders, forward only.
    Players: exact fp32 sampling + ray-bender pre-pass, field on the tensor cores over the non-empty tiles; beside it the same
    frame with the players' field on the fp32 CUDA-core kernel (PE_TC_PREPASS=0)."""
    import scenes
    from helpers import INPUT_KEYS
    from gpu_common import build_composer
    scene = scenes.scene_tennis(seed=13, height=144, width=256, stride=1, lead=(1, 1, 1), dense=dense)

    def timed(precision, reps=5):
        _, _, _, comp, dev = build_composer(scene, precision, device=device)
        call = [dev[k] for k in INPUT_KEYS]
        with torch.no_grad():
            comp(*call, False)
            torch.cuda.synchronize()
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(reps):
                comp(*call, False)
            e0.record()
            torch.cuda.synchronize()
        return s0.elapsed_time(e0) / reps

    out = {"workload": f"cfg3 Tennis{' (dense)' if dense else ''}, eval forward, 256x144 rays, court P=4 + 2 players P=32 with ray benders"}
    for precision in ("mixed", "fp16x3", "fp16"):
        out[f"{precision}_ms"] = timed(precision)
    os.environ["PE_TC_PREPASS"] = "0"
    try:
        out["players_on_fp32_field_ms"] = timed("fp16x3")
    finally:
        del os.environ["PE_TC_PREPASS"]
    return out


def tennis_four_objects(lead, dirs_of):
    """The shipped Tennis scene (SURVEY section 8, T-frame / T-train): 2 static boxes with 4 samples per ray + 2 players with 32 samples
    per ray and positional ray benders = 72 samples per ray.  ``dirs_of(stride)`` -> (origins, directions, normals) of one strided grid."""
    import numpy as np
    import scenes
    court = scenes.object_cfg([[-30, 30], [-40, 20.585], [-0.5, 0.0]], 4, 5.0, 70.0, 64, 32, scenes.nerf_cfg(), scenes.bender_cfg("zeroed"))
    # the second static object of the shipped config is a thin upright slab (configs/tennis/193_*.yaml:183), posed here at the far end of the court
    wall = scenes.object_cfg([[-30, 30], [0.0, 0.5], [0.0, 30.0]], 4, 5.0, 70.0, 64, 32, scenes.nerf_cfg(), scenes.bender_cfg("zeroed"))
    player = lambda: scenes.object_cfg([[-0.75, 0.75], [-0.5, 0.5], [0.0, 2.15]], 32, 5.0, 70.0, 64, 32, scenes.nerf_cfg(), scenes.bender_cfg("positional"))
    config = scenes.scene_config([court, wall, player(), player()], 2, [1, 1, 1, 1], True)
    parts = [dirs_of(st) for st in (4, 8)]
    orig, norm = parts[0][0], parts[0][2]
    dirs = torch.cat([p[1] for p in parts], dim=-2)
    p1 = np.linalg.inv(scenes.homogeneous(scenes.rot_z(0.3), [2.0, -11.0, 0.01]))
    p2 = np.linalg.inv(scenes.homogeneous(scenes.rot_z(-0.2), [-2.0, 11.0, 0.01]))
    pw = np.linalg.inv(scenes.homogeneous(np.eye(3), [0.0, 20.585, 0.0]))
    inputs = scenes.build_inputs(17, config, lead, orig, dirs, norm, [np.eye(4), pw, p1, p2])
    return config, scenes.scene_state(17, config), inputs


def t_train_scene(B: int):
    """The T-train batch (see t_train_report): lead dimensions (B, 4, 1), 5 120 patch rays per image on the near player."""
    import numpy as np
    import scenes
    H, W = 288, 512
    focal = 1700.0 * 0.51417 * 0.5 * (W / 256.0)
    c2w = scenes.tennis_camera()
    # pixel of the near player's centre: the patches the reference's sampler draws are weighted towards the objects' boxes
    pc = c2w[:3, :3].T @ (np.array([2.0, -11.0, 1.0]) - c2w[:3, 3])
    col_c, row_c = W / 2 + focal * pc[0] / -pc[2], H / 2 - focal * pc[1] / -pc[2]
    lead = (B, 4, 1)

    def dirs_of(stride):
        o, d, n = scenes.camera_rays(lead, H, W, focal, c2w, stride)
        gh, gw, side = H // stride, W // stride, 256 // stride
        r0 = int(min(max(row_c / stride - side / 2, 0), gh - side))
        c0 = int(min(max(col_c / stride - side / 2, 0), gw - side))
        d = d.reshape(lead + (gh, gw, 3))[..., r0:r0 + side, c0:c0 + side, :].reshape(lead + (side * side, 3)).contiguous()
        return o, d, n

    return tennis_four_objects(lead, dirs_of), lead


def t_train_report(device, batches=(1, 8), eager=True):
    """What train.py differentiates per step on the shipped Tennis config ("T-train"): B x 4 observations = 4 B images
    of 288x512, per image a 64x64 patch of the stride-4 grid + a 32x32 patch of the stride-8 grid (5 120 rays, both patches on the near
    player), 4 object instances (72 samples per ray), train-mode BatchNorm, perturb=True (stratified jitter + raw-alpha noise), every
    parameter and differentiable input requiring a gradient.  B = 1 is one nn.DataParallel replica of the reference's 8-GPU run
    (train.py:61), B = 8 the whole batch on ONE B200.  Beside it the UPSTREAM composer in eager PyTorch on this GPU (B = 1)."""
    import scenes
    from helpers import INPUT_KEYS
    from gpu_common import build_composer

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            fn()
        e0.record()
        torch.cuda.synchronize()
        return s0.elapsed_time(e0) / reps

    out = []
    for B in batches:
        scene, lead = t_train_scene(B)
        config, state, inputs, comp, dev = build_composer(scene, "mixed", device=device, training=True)
        comp.allow_forward_without_grad = False
        dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
        call = [dev[k] for k in INPUT_KEYS]
        rays = dev["ray_directions"].size(-2)
        cot = torch.randn(lead + (rays, 192), device=device)

        def loss_of(res):
            return (res["global"]["integrated_features"] * cot).sum() + res["global"]["opacity"].sum()

        def fwd_bwd():
            comp.zero_grad(set_to_none=True)
            loss_of(comp(*call, True)["coarse"]).backward()

        comp.return_raw_alphas = True
        with torch.no_grad():
            res = comp(*call, False)["coarse"]
        comp.return_raw_alphas = False
        inbox = sum(int((res[f"object_{k}"]["raw_alphas"] != float(m["empty_space_alpha"])).sum().item())
                    for k, m in enumerate(config["model"]["object_models"]))
        del res
        row = {"workload": f"T-train: Tennis, {4 * B} images x 5120 patch rays (64x64 @ stride 4 + 32x32 @ stride 8), 4 objects, 72 samples per ray, "
                           "train-mode BatchNorm, perturb=True, forward + backward", "batch": B, "rays": 4 * B * rays,
               "sample_slots": 4 * B * rays * 72, "in_box_samples": inbox, "fwd_bwd_ms": timed(fwd_bwd, 3)}
        row["in_box_samples_per_s_fwd_bwd"] = inbox / (row["fwd_bwd_ms"] / 1e3)
        del comp
        if eager and B == 1 and upstream_composer_class(cpu=False) is not None:
            prev = torch.backends.cuda.matmul.allow_tf32
            try:
                with torch.device(device):
                    up = upstream_composer(config, state, device)
                    up.train()

                    def up_step():
                        up.zero_grad(set_to_none=True)
                        for t in call:
                            t.grad = None
                        loss_of(up(*call, True)["coarse"]).backward()

                    for tf32 in (False, True):
                        torch.backends.cuda.matmul.allow_tf32 = tf32
                        row[f"upstream_eager_{'tf32' if tf32 else 'fp32'}_ms"] = timed(up_step, 2)
                    del up
            except Exception as exc:      # noqa: BLE001  (a secondary figure: never fail the bench line for it)
                row["upstream_eager_error"] = f"{type(exc).__name__}: {str(exc)[:200]}"
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
        torch.cuda.empty_cache()
        out.append(row)
    return out


def t_frame_report(device):
    """What play.py renders per frame ("T-frame"): the Tennis full frame 288x512 sampled on the strided grids of the
    multiresolution decoder (strides 4 and 8: 9216 + 2304 = 11 520 rays), 4 object instances (2 static boxes with 4 samples per ray, 2
    players with 32 samples per ray and positional ray benders: 72 samples per ray, 829 440 sample slots), eval mode, the composed features
    handed off as the decoder's per-stride CHW grids (written by the compositor).  ONE composer call per frame (the reference: 12 chunks
    of 1000 rays)."""
    import numpy as np
    import scenes
    from helpers import INPUT_KEYS
    from gpu_common import build_composer
    from playableenvironments_b200.utils.lib_3d.ray_helper import RayHelper
    H, W, strides = 288, 512, [4, 8]
    lead = (1, 1, 1)
    focal = 1700.0 * 0.51417 * 0.5 * (W / 256.0)
    scene = tennis_four_objects(lead, lambda st: scenes.camera_rays(lead, H, W, focal, scenes.tennis_camera(), st))
    dirs = scene[2]["ray_directions"]
    out = {"workload": "T-frame: Tennis 288x512 on the stride-4 + stride-8 grids = 11520 rays, 2 static objects (P=4) + 2 players (P=32, ray benders), "
                       "eval forward, the decoder's per-stride CHW grids written by the compositor (PeHandoff), one composer call", "rays": int(dirs.size(-2)), "sample_slots": int(dirs.size(-2)) * 72}
    for precision in ("mixed", "fp16x3", "fp16"):
        _, _, _, comp, dev = build_composer(scene, precision, device=device)
        call = [dev[k] for k in INPUT_KEYS]

        only_global = [False]

        def frame():
            with torch.no_grad():      # the compositor writes the decoder's per-stride CHW grids itself (PeHandoff)
                return comp(*call, False, handoff=(strides, (H, W), [64, 128]), global_only=only_global[0])["coarse"]["global"]["feature_grids"]

        frame()
        torch.cuda.synchronize()
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(10):
            grids = frame()
        e0.record()
        torch.cuda.synchronize()
        out[f"{precision}_ms"] = s0.elapsed_time(e0) / 10
        if precision == "mixed":
            # what the decoder path consumes (environment_model_multiresolution_backpropagated_decoder.py:84): the composed scene only
            only_global[0] = True
            frame()
            torch.cuda.synchronize()
            s0.record()
            for _ in range(10):
                grids_g = frame()
            e0.record()
            torch.cuda.synchronize()
            out["mixed_global_only_ms"] = s0.elapsed_time(e0) / 10
            out["global_only_equals_full"] = bool(all(torch.equal(a, b) for a, b in zip(grids_g, grids)))
            # the same frame replayed from ONE CUDA graph (the composer call allocates nothing, syncs nothing and sizes its tile lists on
            # the device, so the whole ~30-launch sequence is capturable): what an interactive caller (play.py) should do per frame
            for tag, flag in (("mixed_cuda_graph", False), ("mixed_global_only_cuda_graph", True)):
                only_global[0] = flag
                try:
                    side = torch.cuda.Stream(device=device)
                    side.wait_stream(torch.cuda.current_stream(device))
                    with torch.cuda.stream(side):
                        for _ in range(2):
                            frame()
                    torch.cuda.current_stream(device).wait_stream(side)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        captured = frame()
                    graph.replay()
                    torch.cuda.synchronize()
                    same = all(torch.equal(a, b) for a, b in zip(captured, grids))
                    s0.record()
                    for _ in range(20):
                        graph.replay()
                    e0.record()
                    torch.cuda.synchronize()
                    out[f"{tag}_ms"] = s0.elapsed_time(e0) / 20
                    out["cuda_graph_equals_eager"] = bool(same) and out.get("cuda_graph_equals_eager", True)
                    del graph
                except Exception as exc:      # noqa: BLE001  (a diagnostic figure: never fail the bench line for it)
                    out["cuda_graph_error"] = str(exc)[:200]
            only_global[0] = False
    out["grids"] = [list(g.shape) for g in grids]
    return out


def run_b200(args):
    import torch.distributed as dist
    from helpers import INPUT_KEYS
    from gpu_common import build_composer
    from playableenvironments_b200.model import render

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    scene = build_scene()
    config, state, inputs, comp, dev = build_composer(scene, args.precision, device=device)
    call_args = [dev[k] for k in INPUT_KEYS]
    rays = dev["ray_directions"].size(-2)
    samples_per_rank = rays * POSITIONS
    F = 192
    gathered = torch.empty((world, rays, F), dtype=torch.float32, device=device) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)     # > 126 MB L2

    from playableenvironments_b200 import sharding
    side = torch.cuda.Stream(device=device)

    fused = world > 1 and args.gather == "fused"
    pg, fused_note = None, None
    if fused:
        # symmetric memory needs P2P-capable GPUs under one driver; if the rendezvous fails on ANY rank, all ranks fall back to NCCL
        try:
            pg = sharding.PeerGather(rays, F, device)
            ok = torch.ones(1, device=device)
        except Exception as exc:      # noqa: BLE001
            fused_note = f"PeerGather unavailable ({type(exc).__name__}: {str(exc)[:120]}): NCCL all-gather instead"
            ok = torch.zeros(1, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            fused, pg = False, None
            fused_note = fused_note or "PeerGather unavailable on another rank: NCCL all-gather instead"

    def step():
        with torch.no_grad():
            if fused:
                # the all-gather (the single collective of the path) is FUSED into the render kernel: every ray's features are stored
                # to all ranks' grids over NVLink as they are produced (sharding.PeerGather); a device-side barrier orders the readers
                comp(*call_args, False, peer_features=pg.destinations())
                return pg.sync()[rank]
            if world > 1:
                # --gather nccl: the frame renders in ray chunks; chunk i's grid is all-gathered by NCCL on a side stream while chunk
                # i+1 renders
                return sharding.render_pipelined(comp, *call_args, False, chunks=args.chunks, gathered=gathered, side_stream=side)
            res = comp(*call_args, False)
        return res["coarse"]["global"]["integrated_features"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    render.take_launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()                       # L2 flush between timed iterations (outside the per-step events)
            starts[i].record()
            step()
            ends[i].record()
        barrier()
        wall = time.perf_counter() - wall0
    launches = render.take_launch_count()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    value = world * samples_per_rank * args.steps / total_s

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----
    host_in = {k: v.contiguous().pin_memory() for k, v in inputs.items()}
    host_out = torch.empty((rays, F), dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    d2h = host_out.numel() * 4

    def e2e_step():
        d = [host_in[k].to(device, non_blocking=True) for k in INPUT_KEYS]
        with torch.no_grad():      # D2H of chunk i (and its all-gather) overlap the render of chunk i+1
            sharding.render_pipelined(comp, *d, False, chunks=args.chunks, gathered=gathered if (world > 1 and not fused) else None,
                                      host_out=host_out, side_stream=side, peer_gather=pg)

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = args.steps
    s_ev.record()
    for _ in range(e2e_steps):
        e2e_step()
    e_ev.record()
    barrier()
    e2e_ms = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * samples_per_rank * e2e_steps / (float(e2e_ms.item()) / 1e3)

    # ---- the other tensor-core modes on the same frame; live parity of the timed mode against the UPSTREAM reference's output on every
    # 16th ray of this very frame (tests/golden/cfg2_subset.npz, written by tests/golden/make_golden_fullsize.py) ----
    modes, parity = {}, None
    if rank == 0 and world == 1 and not args.quick:
        import numpy as np
        from helpers import load_golden
        feats_main = step().clone()
        g = load_golden("cfg2_subset")
        stride = int(g["stride"])
        # rays whose last-sample raw alpha sits on the opacity step (alpha jumps 0 -> 1 over a 1e10 interval,
        # object_composer.py:172,197) are not resolvable below fp32: counted, and excluded from the bound
        stable = torch.from_numpy(np.abs(g["raw_alpha_last"].reshape(-1)) > 4e-3).to(device)
        ref = torch.from_numpy(g["integrated_features"].reshape(-1, F)).to(device).double()
        got = feats_main.reshape(rays, F)[::stride].double()
        diff = (got - ref)[stable]
        parity = {"reference": "upstream ObjectComposer on every 16th ray of this frame (4096 rays, tests/golden/cfg2_subset.npz)",
                  "rel_l2": float(diff.norm() / ref[stable].norm()), "max_over_scale": float(diff.abs().max() / ref.abs().max()),
                  "max_over_scale_all_rays": float((got - ref).abs().max() / ref.abs().max()),
                  "frac_of_values_within_1e-3_of_scale": float((diff.abs() <= 1e-3 * ref.abs().max()).double().mean()),
                  "rays_on_opacity_step_excluded": float(1.0 - stable.float().mean())}
        for other in ("fp16", "mixed", "fp16x2", "fp16x3"):
            if other == args.precision:
                continue
            comp.precision = other
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_other = max(3, args.steps // 4)
            s0.record()
            for _ in range(n_other):
                out_other = step()
            e0.record()
            torch.cuda.synchronize()
            ms = s0.elapsed_time(e0) / n_other
            d_o = (out_other.reshape(rays, F)[::stride].double() - ref)[stable]
            modes[other] = {"ms_per_step": ms, "value": samples_per_rank / (ms / 1e3),
                            "roofline_frac": samples_per_rank * FLOP_PER_SAMPLE / (ms / 1e3) / 1e12 / measured_peaks()["burst"],
                            "max_over_scale_vs_reference": float(d_o.abs().max() / ref.abs().max())}
        comp.precision = args.precision

    train_multi = None
    if world > 1 and not args.quick:
        train_multi = train_step_report(device, True, world=world)       # every rank takes part (collective inside)
    if rank == 0:
        peaks = measured_peaks()
        ms_per_step = total_s * 1e3 / args.steps
        # dominant kernel = the fused field kernel; the step is that kernel plus two tiny style prologues (share > 99.9 %, profiles/)
        achieved_tflops = samples_per_rank * FLOP_PER_SAMPLE / (ms_per_step / 1e3) / 1e12
        line = {
            "metric": "ray-samples/sec (style-MLP ray-march, fwd)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"mixed": "f16 operands (activation-aware weight rounding, weights hi+lo on trunk layers L6-L7), f32 accumulate",
                                           "fp16": "f16 operands, f32 accumulate", "fp16x2": "f16 operands (weights hi+lo), f32 accumulate",
                                           "fp16x3": "f16 operands (weights and activations hi+lo), f32 accumulate", "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "precision": args.precision, "frames_per_gpu_per_step": 1, "l2": "flushed between steps (256 MB memset)",
                       "frames_per_s": world * args.steps / total_s, "collective_note": fused_note, "collective": ("none" if world == 1 else "all-gather of the feature grid fused into the render kernel (P2P stores over NVLink into symmetric memory, sharding.PeerGather) + one signal-pad barrier"
                                      if fused else f"NCCL all_gather(feature grid) per ray chunk, {args.chunks} chunks, on a side stream")},
            "roofline": {"bound": "tensor", "achieved": achieved_tflops, "peak": peaks["burst"], "unit": "TFLOP/s",
                         "frac": achieved_tflops / peaks["burst"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one pe_field_tc_kernel launch on this workload
                         # (profiles/r1c_final_summary.md; fp16: 3.4 MB + 146.8 MB, fp16x2: 4.1 MB + 147.7 MB)
                         "traffic": {"fp16": 150208768, "fp16x2": 151788032, "mixed": TRAFFIC_MIXED}.get(args.precision), "peak_source": peaks["source"] + " bf16 burst",
                         "frac_of_sustained": achieved_tflops / peaks["sustained"],
                         "tensor_passes": {"fp16": 1, "fp16x2": 2, "fp16x3": 3, "fp32": 0, "mixed": "2 on L6-L7, 1 on L0-L5 and the head (activation-aware fp16 weight stream)"}[args.precision]},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "wall_s": wall,
        }
        if world == 1 and not args.quick:
            line["train_step"] = [train_step_report(device, False), train_step_report(device, True),
                                  train_step_report(device, False, "fp32"), train_step_report(device, True, "fp32")]
            line["eval_frame"] = [eval_frame_report(device, False), eval_frame_report(device, True)]
            line["t_frame"] = t_frame_report(device)
            line["t_train"] = t_train_report(device)
            line["gpu_eager_baseline"] = gpu_eager_reference(device)
        if train_multi:
            line["train_step"] = [train_multi]
        if modes:
            line["other_modes"] = modes
        if parity:
            line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            # in its own process: the CPU shims of the upstream code (no-op .cuda()) must not leak into this one
            import subprocess
            proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                                  capture_output=True, text=True, timeout=900)
            ref_line = next((json.loads(l) for l in proc.stdout.splitlines() if l.startswith("{")), None)
            if ref_line is None:
                raise RuntimeError("cpu baseline failed: " + proc.stderr[-400:])
            line["cpu_baseline"] = ref_line["cpu_baseline"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PE_PRECISION", "mixed"), choices=["mixed", "fp16", "fp16x2", "fp16x3", "fp32"])
    ap.add_argument("--chunks", type=int, default=4, help="ray chunks per frame of the pipelined render (N > 1 and the e2e figure)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="N > 1: all-gather fused into the render kernel, or NCCL per ray chunk")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the other precision modes and the live parity check")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
