import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import bench
torch.cuda.set_device(0)
d = torch.device("cuda", 0)
for dense in (False, True):
    r = bench.eval_frame_report(d, dense)
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k != "workload"}), flush=True)
t = bench.t_frame_report(d)
print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items() if k.endswith("_ms")}))
