"""Driver for compute-sanitizer (memcheck) over the last paths of round 2: the compositor's decoder hand-off (PeHandoff) and the
global_only call on a frame made of two strided grids, the ray bender with split TMEM accumulators, the per-object stream fork / join,
and a train step with the exact backward tile counts forced on."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch  # noqa: E402
import scenes  # noqa: E402
from helpers import INPUT_KEYS  # noqa: E402
from gpu_common import build_composer  # noqa: E402

os.environ["PE_BWD_TILE_COUNTS"] = "1"
H, W, strides, channels = 32, 64, [4, 8], [64, 128]
lead = (1, 2, 1)
parts = [scenes.camera_rays(lead, H, W, 60.0, scenes.tennis_camera(), st) for st in strides]
base = scenes.scene_tennis(seed=13, height=H, width=W, stride=4, lead=lead)
inputs = dict(base[2])
inputs["ray_directions"] = torch.cat([p[1] for p in parts], dim=-2)
for precision in ("mixed", "fp16"):
    _, _, _, comp, dev = build_composer((base[0], base[1], inputs), precision)
    call = [dev[k] for k in INPUT_KEYS]
    with torch.no_grad():
        comp(*call, False, handoff=(strides, (H, W), channels))
        comp(*call, False, handoff=(strides, (H, W), channels), global_only=True)
        comp(*call, False, global_only=True)
    torch.cuda.synchronize()
    print("hand-off", precision, "ok", flush=True)
for name in ("tennis_dense", "minecraft_small"):
    _, _, _, comp, dev = build_composer(name, "mixed", training=True)
    comp.allow_forward_without_grad = False
    dev = {k: (v.clone().requires_grad_(True) if k in scenes.GRAD_INPUT_KEYS else v) for k, v in dev.items()}
    res = comp(*[dev[k] for k in INPUT_KEYS], True)["coarse"]
    (res["global"]["integrated_features"].sum() + res["global"]["opacity"].sum()).backward()
    torch.cuda.synchronize()
    print("train step", name, "ok", flush=True)
