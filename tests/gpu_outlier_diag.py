import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE, os.path.join(_HERE, "golden")):
    sys.path.insert(0, _p)
import numpy as np, torch, scenes
from helpers import load_golden
from gpu_common import build_composer, run_composer
os.environ["PE_TC_AWARE_MASK"] = sys.argv[1] if len(sys.argv) > 1 else "0x000"
g = load_golden("static_small")
def raws(precision):
    _, _, _, comp, dev = build_composer("static_small", precision)
    comp.return_raw_alphas = True
    r = run_composer(comp, dev)["coarse"]["object_0"]
    return r["weights"].reshape(-1, 128).cpu().numpy(), r["raw_alphas"].reshape(-1, 128).cpu().numpy()
w32, a32 = raws("fp32")
wm, am = raws("mixed")
wref = g["coarse/object_0/weights"].reshape(-1, 128)
err = np.abs(wm - wref); scale = wref.max()
print("scale", scale, "max err/scale", err.max() / scale, "fp32 path err", np.abs(w32 - wref).max() / scale)
order = np.argsort(err.reshape(-1))[::-1][:8]
for o in order:
    r, p = divmod(int(o), 128)
    print(f"ray {r} sample {p}: err/scale {err[r,p]/scale:.2e} w_ref {wref[r,p]:.4f} raw32 {a32[r,p]:.4f} raw_mixed {am[r,p]:.4f} d_raw {am[r,p]-a32[r,p]:+.2e}  raw err of neighbours {np.abs(am[r,max(0,p-3):p+1]-a32[r,max(0,p-3):p+1]).round(5)}")
d = np.abs(am - a32)
print("raw alpha error: mean %.2e  p99 %.2e  max %.2e ; |raw| mean %.2f" % (d.mean(), np.quantile(d, 0.99), d.max(), np.abs(a32).mean()))
print("rays with err > 8e-4:", np.unique(np.argwhere(err / scale > 8e-4)[:, 0]))
