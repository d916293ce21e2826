"""Diagnostic: a render on explicit ray parameters equal to the stratified ones must reproduce the stratified render, forward and backward."""
import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE, os.path.join(_HERE, "golden")):
    sys.path.insert(0, _p)
import numpy as np, torch
import scenes
from helpers import INPUT_KEYS
from gpu_common import build_composer
from playableenvironments_b200 import _cabi
from playableenvironments_b200.model import render

def run(name, precision, bwd_tc, explicit):
    os.environ["PE_BWD_TC"] = bwd_tc
    _, _, _, comp, dev = build_composer(name, precision)
    helper = comp.object_id_helper
    K = helper.objects_count
    descs = comp._descs(False)
    args = dict(ray_origins=dev["ray_origins"], ray_directions=dev["ray_directions"].clone().requires_grad_(True), w2o=dev["transformation_matrix_w2o"],
                style=dev["style"], deformation=dev["deformation"], object_in_scene=dev["object_in_scene"])
    common = (descs, helper.static_objects_count, args["ray_origins"], args["ray_directions"], args["w2o"], args["style"], args["deformation"],
              args["object_in_scene"], False, False, True, False, _cabi.PRECISIONS[precision])
    sample_t = None
    if explicit:
        with torch.no_grad():
            r = render.render_scene(*common, return_samples=True)
        sample_t = [r[f"object_{k}"]["positions_t"].clone() for k in range(K)]
    models = [comp.object_models_coarse[helper.model_idx_by_object_idx(k)] for k in range(K)]
    res = render.render_scene(*common, models=models, sample_t=sample_t)
    g = torch.Generator(device="cuda").manual_seed(1)
    loss = sum((res[n][key] * torch.randn(res[n][key].shape, device="cuda", generator=g)).sum() for n in res for key in ("integrated_features", "opacity", "depth"))
    loss.backward()
    torch.cuda.synchronize()
    out = {k: p.grad.cpu().numpy() for k, p in comp.named_parameters() if p.grad is not None}
    out["in/ray_directions"] = args["ray_directions"].grad.cpu().numpy()
    out["fwd/features"] = res["global"]["integrated_features"].detach().cpu().numpy()
    return out

for name in sys.argv[1:] or ["static_small"]:
    for prec, tc in (("fp32", "0"), ("fp16x3", "1")):
        a, b = run(name, prec, tc, False), run(name, prec, tc, True)
        errs = {k: float(np.abs(b[k] - a[k]).max() / max(np.abs(a[k]).max(), 1e-12)) for k in a}
        top = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
        print(name, prec, "PE_BWD_TC=" + tc, [(k[-50:], float(f"{v:.3g}")) for k, v in top], flush=True)
