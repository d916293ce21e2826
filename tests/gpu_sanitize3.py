"""Driver for compute-sanitizer over the paths changed at the end of round 2: merged ordering of the composed list (ordered lists, lists with
samples masked by fix_object_overlaps), batched dot products of the compositing backward, NULL cotangents for unused outputs, the segmented
absmax launch.  Usage: compute-sanitizer --tool memcheck python tests/gpu_sanitize3.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch  # noqa: E402
import scenes  # noqa: E402
from helpers import INPUT_KEYS  # noqa: E402
from gpu_common import build_composer, run_composer  # noqa: E402

for name in ("tennis_dense", "tennis_small", "minecraft_small", "minecraft_absent", "toy_world"):
    _, _, _, comp, dev = build_composer(name, "mixed")
    run_composer(comp, dev)
    torch.cuda.synchronize()
    print("eval", name, "ok", flush=True)
for name, training in (("tennis_dense", True), ("minecraft_small", False), ("toy_world", False)):
    _, _, _, comp, dev = build_composer(name, "mixed", training=training)
    comp.allow_forward_without_grad = False
    dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
    res = comp(*[dev[k] for k in INPUT_KEYS], False)["coarse"]
    (res["global"]["integrated_features"].sum() + res["global"]["opacity"].sum()).backward()      # per-object cotangents stay NULL
    torch.cuda.synchronize()
    res = comp(*[dev[k] for k in INPUT_KEYS], False)["coarse"]
    sum(v["integrated_features"].sum() + v["depth"].sum() + v["weights"].sum() for v in res.values()).backward()
    torch.cuda.synchronize()
    print("backward", name, "ok", flush=True)
