"""Driver for compute-sanitizer over the paths added in the second half of round 2: the fine pass (forward + backward), the Hutchinson
divergence, the backward of forward_expected_positions, the activation-aware weight packing with the mixed eval forward, the train-mode
trunk cache.  Usage: compute-sanitizer --tool memcheck python tests/gpu_sanitize2.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch  # noqa: E402
import scenes  # noqa: E402
from helpers import INPUT_KEYS  # noqa: E402
from gpu_common import build_composer, run_composer  # noqa: E402

# 1. aware packing + mixed eval forward, single object (folded head) and multi-object
for name in ("static_small", "tennis_dense"):
    _, _, _, comp, dev = build_composer(name, "mixed")
    run_composer(comp, dev)
torch.cuda.synchronize()
print("eval ok", flush=True)
# 2. fine pass forward + backward
_, _, _, comp, dev = build_composer(scenes.FINE_SCENES["tennis_fine"](), "mixed")
comp.allow_forward_without_grad = False
dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
res = comp(*[dev[k] for k in INPUT_KEYS], False)
(res["fine"]["global"]["integrated_features"].sum() + res["coarse"]["global"]["opacity"].sum()).backward()
torch.cuda.synchronize()
print("fine ok", flush=True)
# 3. train mode (trunk cache) with the divergence, forward + backward
config, _, inputs, comp, dev = build_composer("tennis_dense", "mixed", training=True)
comp.compute_divergence = True
comp.allow_forward_without_grad = False
res = comp(*[dev[k] for k in INPUT_KEYS], False)
(res["coarse"]["global"]["integrated_features"].sum() + res["coarse"]["global"]["integrated_divergence"].sum()).backward()
torch.cuda.synchronize()
print("train + divergence ok", flush=True)
# 4. expected positions backward
from make_golden_expected import object_inputs  # noqa: E402
_, _, inputs, comp, _ = build_composer("tennis_dense", "fp16x3")
comp.allow_forward_without_grad = False
args = [t.cuda() for t in object_inputs(inputs, 1)]
args[1] = args[1].clone().requires_grad_(True)
exp, opacity = comp.forward_expected_positions(*args, 1, False)["coarse"]
(exp.sum() + opacity.sum()).backward()
torch.cuda.synchronize()
print("expected positions ok", flush=True)
