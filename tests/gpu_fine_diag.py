"""Diagnostic: parameter gradients of static_fine, tensor-core backward vs exact fp32 backward (same explicit ray parameters)."""
import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE, os.path.join(_HERE, "golden")):
    sys.path.insert(0, _p)
import numpy as np, torch
import scenes
from helpers import INPUT_KEYS, load_golden
from test_gpu_fine import _build, _feed_golden_coarse_weights, _fine_loss

def grads(precision, bwd_tc, name="static_fine"):
    os.environ["PE_BWD_TC"] = bwd_tc
    golden = load_golden(f"{name}_grad")
    _, _, _, comp, dev = _build(name, precision)
    _feed_golden_coarse_weights(comp, load_golden(name), "cuda")
    comp.allow_forward_without_grad = False
    dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
    res = comp(*[dev[k] for k in INPUT_KEYS], False)
    loss = _fine_loss(res, [str(k) for k in golden["loss_keys"]])
    loss.backward()
    torch.cuda.synchronize()
    out = {k: p.grad.cpu().numpy() for k, p in comp.named_parameters() if p.grad is not None}
    out.update({"in/" + k: dev[k].grad.cpu().numpy() for k in scenes.GRAD_INPUT_KEYS if dev[k].grad is not None})
    return out

name = sys.argv[1] if len(sys.argv) > 1 else "static_fine"
a = grads("fp32", "0", name)
for prec, tc in (("fp16x3", "1"), ("fp16x3", "0"), ("fp32", "1")):
    b = grads(prec, tc, name)
    errs = {k: float(np.abs(b[k] - a[k]).max() / max(np.abs(a[k]).max(), 1e-12)) for k in a}
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    print(prec, "PE_BWD_TC=" + tc, [(k[-60:], round(v, 4)) for k, v in top], flush=True)
