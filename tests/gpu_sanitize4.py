"""Driver for compute-sanitizer --tool initcheck: one train-mode forward + backward of tennis_dense (mixed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import scenes
from helpers import INPUT_KEYS
from gpu_common import build_composer
os.environ.setdefault("PYTORCH_NO_CUDA_MEMORY_CACHING", "1")
config, state, inputs, comp, dev = build_composer("tennis_dense", "mixed", training=True)
comp.allow_forward_without_grad = False
dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
res = comp(*[dev[k] for k in INPUT_KEYS], False)["coarse"]
scenes.grad_loss(res, ["global/integrated_features", "global/opacity", "global/depth", "object_1/opacity"]).backward()
torch.cuda.synchronize()
print("done")
