"""Diagnostic: is a train-mode forward + backward of tennis_dense (mixed) reproducible run to run?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import scenes
from helpers import INPUT_KEYS
from gpu_common import build_composer


def run(poison=None):
    if poison is not None:
        # fill the caching allocator's free memory with a byte pattern: any read of memory the call did not write itself shows up
        junk = torch.empty(3 << 30, dtype=torch.uint8, device="cuda").fill_(poison)
        del junk
    config, state, inputs, comp, dev = build_composer("tennis_dense", "mixed", training=True)
    comp.allow_forward_without_grad = False
    dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
    res = comp(*[dev[k] for k in INPUT_KEYS], False)["coarse"]
    scenes.grad_loss(res, ["global/integrated_features", "global/opacity", "global/depth", "object_1/opacity"]).backward()
    torch.cuda.synchronize()
    out = {"fwd/" + n + "/" + k: v.detach().clone() for n, r in res.items() for k, v in r.items() if torch.is_tensor(v)}
    out.update({"gin/" + k: dev[k].grad.clone() for k in scenes.GRAD_INPUT_KEYS if dev[k].grad is not None})
    out.update({"gpar/" + k: p.grad.clone() for k, p in comp.named_parameters() if p.grad is not None})
    return out


a = run()
for i, poison in enumerate((None, 0x00, 0xFF, 0x3C)):
    b = run(poison)
    worst = sorted(((float((a[k] - b[k]).abs().max() / a[k].abs().max().clamp_min(1e-30)), k) for k in a if "divergence" not in k and "disparity" not in k), reverse=True)[:4]
    print(i, worst)
