"""Diagnostic: T-train replica step, parameter-gradient errors against the upstream composer on the GPU for (a) the default tensor-core
backward, (b) the exact fp32 backward (PE_BWD_TC=0), and (c) the upstream composer in float64 against itself in float32 (conditioning)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import bench, scenes
from helpers import INPUT_KEYS
from gpu_common import build_composer

device = torch.device("cuda", 0)
scene, lead = bench.t_train_scene(1)
torch.manual_seed(0)
cot_f = torch.randn(lead + (5120, 192), device=device)
cot_o = torch.randn(lead + (5120,), device=device)


def loss_of(res, dt=torch.float32):
    g = res["coarse"]["global"]
    return (g["integrated_features"] * cot_f.to(dt)).sum() + (g["opacity"] * cot_o.to(dt)).sum() + (res["coarse"]["object_2"]["opacity"] * cot_o.to(dt)).sum()


def ours(env):
    for k, v in env.items():
        os.environ[k] = v
    config, state, inputs, comp, dev = build_composer(scene, "mixed", device=device, training=True)
    comp.allow_forward_without_grad = False
    call = [dev[k] for k in INPUT_KEYS]
    loss_of(comp(*call, False)).backward()
    for k in env:
        os.environ.pop(k)
    return {k: p.grad.double() for k, p in comp.named_parameters() if p.grad is not None}, config, state, dev


def upstream(config, state, dev, dt):
    with torch.device(device):
        up = bench.upstream_composer(config, state, device)
        up.train()
        # (the upstream divergence pass differentiates w.r.t. the sample positions: they must carry a graph, i.e. the rays require grad)
        call = [(dev[k].clone().requires_grad_(k in scenes.GRAD_INPUT_KEYS)) if dev[k].is_floating_point() else dev[k] for k in INPUT_KEYS]
        loss_of(up(*call, False), dt).backward()
    return {k: p.grad.double() for k, p in up.named_parameters() if p.grad is not None}


def port(config, state, dev, dt):
    """The oracle port of the same graph (pinned against upstream by tests/golden) -- the upstream modules hard-code float32, the port
    follows the default dtype."""
    from oracle import render_oracle as O
    torch.set_default_dtype(dt)
    try:
        with torch.device(device):
            sd = {k: ((v.to(device).to(dt).requires_grad_("running_" not in k)) if v.is_floating_point() else v.to(device)) for k, v in state.items()}
            call = [(dev[k].detach().to(dt)) if dev[k].is_floating_point() else dev[k] for k in INPUT_KEYS]
            loss_of(O.composer_forward(config, sd, *call, perturb=False, training=True), dt).backward()
    finally:
        torch.set_default_dtype(torch.float32)
    return {k: v.grad.double() for k, v in sd.items() if torch.is_tensor(v) and v.is_floating_point() and v.grad is not None}


def worst(a, b, n=6):
    e = {k: float((a[k] - b[k]).abs().max() / b[k].abs().max().clamp_min(1e-30)) for k in b if float(b[k].abs().max()) > 0}
    return sorted(e.items(), key=lambda kv: -kv[1])[:n]


def l2(a, b):
    out = {}
    for obj in ("0", "1", "2", "3", ""):
        ks = [k for k in b if k.startswith("object_models_coarse." + obj) and k in a]
        num = sum(float(((a[k] - b[k]) ** 2).sum()) for k in ks) ** 0.5
        den = sum(float((b[k] ** 2).sum()) for k in ks) ** 0.5
        out[obj or "all"] = num / max(den, 1e-300)
    return out


tc, config, state, dev = ours({})
fp, _, _, _ = ours({"PE_BWD_TC": "0"})
wlo, _, _, _ = ours({"PE_BWD_CHAIN_WLO": "1"})
tc2, _, _, _ = ours({"PE_TC_BENDER": "0"})
tc3, _, _, _ = ours({"PE_TC_BENDER": "0", "PE_BWD_TC_BENDER": "0"})
u32 = upstream(config, state, dev, torch.float32)
u64 = port(config, state, dev, torch.float64)
p32 = port(config, state, dev, torch.float32)
print(json.dumps({"L2 tc_wlo_vs_up32": l2(wlo, u32), "L2 tc_fp32bender_vs_up32": l2(tc2, u32), "L2 tc_fp32bender_fwd_and_bwd_vs_up32": l2(tc3, u32)}))
print(json.dumps({"L2 tc_vs_up32": l2(tc, u32), "L2 fp32bwd_vs_up32": l2(fp, u32), "L2 up32_vs_64": l2(u32, u64), "L2 tc_vs_64": l2(tc, u64)}))
print(json.dumps({"tc_vs_up32": worst(tc, u32), "fp32bwd_vs_up32": worst(fp, u32), "up32_vs_up64": worst(u32, u64), "port32_vs_up32": worst(p32, u32), "port32_vs_64": worst(p32, u64), "tc_vs_up64": worst(tc, u64),
                  "fp32bwd_vs_up64": worst(fp, u64)}, indent=1))
