"""Diagnostic: what the L2 flush between bench steps costs the cfg2 frame (cold L2 / dirty lines / clocks)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch, scenes
from gpu_common import build_composer, run_composer
scene = scenes.scene_static(seed=12, height=256, width=256, P=128)
_, _, _, comp, dev = build_composer(scene, "mixed")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
small = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
def timed(pre, n=20):
    for _ in range(3): run_composer(comp, dev)
    torch.cuda.synchronize()
    s = [torch.cuda.Event(enable_timing=True) for _ in range(n)]; e = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    for i in range(n):
        pre()
        s[i].record(); run_composer(comp, dev); e[i].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in zip(s, e))
    return round(sum(t) / n, 3), round(t[0], 3), round(t[-1], 3)
print(json.dumps({"no_flush": timed(lambda: None), "flush_256MB_zero": timed(lambda: flush.zero_()),
                  "flush_then_sleep_2ms": timed(lambda: (flush.zero_(), torch.cuda._sleep(4_000_000))),
                  "sleep_2ms_only": timed(lambda: torch.cuda._sleep(4_000_000)),
                  "flush_32MB": timed(lambda: small.zero_()),
                  "flush_read_only": timed(lambda: flush.view(torch.int32).sum())}))
