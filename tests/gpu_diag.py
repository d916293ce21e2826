"""One-shot GPU diagnostic: runs every parity check without stopping at the first failure and prints a report.
Usage (on the GPU box): python tests/gpu_diag.py [--quick]"""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]

import scenes  # noqa: E402
from helpers import INPUT_KEYS, compare, flatten, load_golden, scale_rel_err  # noqa: E402
from gpu_common import build_composer, run_composer  # noqa: E402
from playableenvironments_b200 import _cabi  # noqa: E402

report = {}
ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]
TAG = "_".join(ONLY) if ONLY else "all"


def section(name):
    def deco(fn):
        if ONLY and not any(name.startswith(o) for o in ONLY):
            return fn
        t0 = time.time()
        try:
            torch.cuda.synchronize()
            out = fn()
            torch.cuda.synchronize()
            report[name] = {"ok": True, "result": out, "s": round(time.time() - t0, 2)}
        except Exception as e:  # noqa: BLE001
            report[name] = {"ok": False, "error": repr(e)[:600], "trace": traceback.format_exc()[-1500:]}
        print(name, json.dumps(report[name], default=str)[:3000], flush=True)
        return fn
    return deco


def errs(res, golden, skip=()):
    flat = flatten(res)
    out = {}
    for k, ref in golden.items():
        if not k.startswith("coarse/") or any(s in k for s in skip):
            continue
        if k not in flat:
            out[k] = "missing"
            continue
        out[k] = float("%.3g" % scale_rel_err(flat[k], ref))
    return out


def worst(d):
    vals = [(v if isinstance(v, float) else float("inf"), k) for k, v in d.items()]
    return max(vals) if vals else (0.0, "")


@section("umma_gemm")
def _():
    out = {}
    for n, k in [(256, 64), (256, 256), (128, 256), (192, 128), (16, 16)]:
        g = torch.Generator(device="cpu").manual_seed(n * 1000 + k)
        a = torch.randn(128, k, generator=g).cuda()
        b = torch.randn(n, k, generator=g).cuda()
        bias = torch.randn(n, generator=g).cuda()
        for with_bias in (False, True):
            d = torch.full((128, n), float("nan"), device="cuda")
            _cabi.check(_cabi.lib().pe_debug_umma_gemm(a.data_ptr(), b.data_ptr(), bias.data_ptr() if with_bias else None, d.data_ptr(), n, k,
                                                       torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            ref = a.half().float() @ b.half().float().t() + (bias if with_bias else 0.0)
            out[f"{n}x{k}" + ("+bias" if with_bias else "")] = float("%.3g" % ((d - ref).abs().max() / ref.abs().max()).item())
    return out


for scene_name in ["cfg1", "static_small", "tennis_small", "tennis_dense", "tennis_anneal", "minecraft_small", "minecraft_absent"]:
    @section(f"fp32/{scene_name}")
    def _(scene_name=scene_name):
        _, _, _, comp, dev = build_composer(scene_name, "fp32")
        e = errs(run_composer(comp, dev), load_golden(scene_name))
        return {"worst": worst(e), "all": e}

for prec in ["fp16", "fp16x2", "fp16x3"]:
    for scene_name in ["static_small", "tennis_small", "minecraft_small", "minecraft_absent"]:
        @section(f"{prec}/{scene_name}")
        def _(scene_name=scene_name, prec=prec):
            _, _, _, comp, dev = build_composer(scene_name, prec)
            e = errs(run_composer(comp, dev), load_golden(scene_name))
            return {"worst": worst(e), "all": e}

for scene_name in ["cfg1", "tennis_dense", "minecraft_small"]:
    @section(f"perturb_fp32/{scene_name}")
    def _(scene_name=scene_name):
        config, _, inputs, comp, dev = build_composer(scene_name, "fp32")
        rand, noise = scenes.perturbation_tensors(7, config, inputs)
        rand = [r.cuda() for r in rand]
        noise = {k: v.cuda() for k, v in noise.items()}
        e = errs(run_composer(comp, dev, perturb=True, rand=rand, noise=noise), load_golden(scene_name + "_perturb"))
        return {"worst": worst(e), "all": e}

for scene_name in ["cfg1", "static_small", "tennis_dense"]:
    @section(f"train_fp32/{scene_name}")
    def _(scene_name=scene_name):
        _, _, _, comp, dev = build_composer(scene_name, "fp32", training=True)
        golden = load_golden(scene_name + "_train")
        e = errs(run_composer(comp, dev), golden, skip=("integrated_divergence",))
        sd = comp.state_dict()
        for k, ref in golden.items():
            if k.startswith("state/"):
                e[k] = float("%.3g" % scale_rel_err(sd[k[6:]].cpu().numpy(), ref))
        return {"worst": worst(e), "all": e}


@section("perturb_fp16/static")
def _():
    scene = scenes.scene_static(seed=21, height=8, width=8, P=128)
    config, _, inputs, comp, dev = build_composer(scene, "fp16")
    rand, noise = scenes.perturbation_tensors(7, config, inputs)
    rand = [r.cuda() for r in rand]
    noise = {k: v.cuda() for k, v in noise.items()}
    res16 = run_composer(comp, dev, perturb=True, rand=rand, noise=noise)
    comp.precision = "fp32"
    res32 = run_composer(comp, dev, perturb=True, rand=rand, noise=noise)
    a, b = flatten(res16), flatten(res32)
    e = {k: float("%.3g" % scale_rel_err(a[k], b[k])) for k in b if k.startswith("coarse/global")}
    return e


for P in [1, 4, 16, 32, 48, 100]:
    @section(f"tc_vs_fp32/P{P}")
    def _(P=P):
        scene = scenes.scene_static(seed=30 + P, height=6, width=10, P=P, lead=(1, 2, 1))
        _, _, _, comp, dev = build_composer(scene, "fp16x2")
        r16 = flatten(run_composer(comp, dev))
        comp.precision = "fp32"
        r32 = flatten(run_composer(comp, dev))
        return {k: float("%.3g" % scale_rel_err(r16[k], r32[k])) for k in r32 if k.startswith("coarse/global")}


@section("ops/positional_encoding")
def _():
    from playableenvironments_b200.model.positional_encoder import PositionalEncoder
    from oracle import render_oracle as O
    x = torch.randn(1000, 3)
    enc = PositionalEncoder(3, 10, True).cuda()
    got = enc(x.cuda()).cpu()
    return float((got - O.positional_encoding(x, 10)).abs().max())


@section("ops/rays_and_fold")
def _():
    from playableenvironments_b200.utils.lib_3d.ray_helper import RayHelper
    from oracle import render_oracle as O
    H, W, strides = 32, 64, [4, 8]
    focal = torch.tensor([[40.0, 55.0]])
    c2w = torch.from_numpy(np.stack([scenes.tennis_camera(), scenes.homogeneous(scenes.rot_x(0.3), [1, 2, 3])])).float().reshape(1, 2, 4, 4)
    d, o, p = RayHelper.generate_strided_grid_rays(focal.cuda(), c2w.cuda(), H, W, strides)
    dirs, org, nrm = O.create_camera_rays([1, 2], H, W, focal)
    sd, sp = O.sample_all_rays_strided_grid(dirs, strides)
    ro, rd, _ = O.transform_rays(org, sd, nrm, c2w)
    out = {"dirs": float((d.cpu() - rd).abs().max()), "orig": float((o.cpu() - ro).abs().max()), "pos": float((p.cpu() - sp).abs().max())}
    R = d.size(-2)
    feats = torch.randn(1, 2, R, 192)
    grids = RayHelper.fold_feature_grids(feats.cuda(), strides, (H, W), [64, 128])
    ref = O.decoder_feature_grids(feats, strides, (H, W), [64, 128])
    out["fold"] = max(float((g.cpu() - r).abs().max()) for g, r in zip(grids, ref))
    return out


@section("ops/field_on_positions")
def _():
    from oracle import render_oracle as O
    config, state, inputs, comp, dev = build_composer("tennis_dense", "fp32")
    model = comp.object_models_coarse[1]
    cfg = config["model"]["object_models"][1]
    g = torch.Generator().manual_seed(5)
    pos = (torch.rand(2, 50, 7, 3, generator=g) - 0.5) * torch.tensor([1.8, 1.2, 2.6]) + torch.tensor([0.0, 0.0, 1.0])
    org = torch.randn(2, 50, 3, generator=g)
    drs = torch.randn(2, 50, 3, generator=g)
    sty = torch.randn(2, 1, 64, generator=g)
    dfm = torch.randn(2, 1, 32, generator=g)
    f, a, d, _ = model(pos.cuda(), org.cuda(), drs.cuda(), sty.cuda(), dfm.cuda())
    sd = {k[len("object_models_coarse.1."):]: v for k, v in state.items() if k.startswith("object_models_coarse.1.")}
    rf, ra, rd = O.ray_bending_style_nerf(sd, cfg, pos, org, drs, sty, dfm, False, False)
    return {"features": scale_rel_err(f.cpu().numpy(), rf.numpy()), "alpha": scale_rel_err(a.cpu().numpy(), ra.numpy()),
            "disp": scale_rel_err(d.cpu().numpy(), rd.numpy())}


print("\n==== SUMMARY ====")
for k, v in report.items():
    if not v["ok"]:
        print(f"FAIL  {k}: {v['error']}")
    else:
        r = v["result"]
        w = r.get("worst") if isinstance(r, dict) and "worst" in r else r
        print(f"ok    {k}: {w}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"diag_{TAG}.json"), "w") as fh:
    json.dump(report, fh, indent=1, default=str)
