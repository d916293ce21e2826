"""Diagnostic: gradients of the tensor-core backward with / without the lo pass of the transposed weight stream (PE_BWD_CHAIN_WLO)
against the exact fp32 backward and the upstream goldens, and the cfg3 train-step times."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import numpy as np, torch
from helpers import compare_grads
from test_gpu_backward import run_backward
import bench
for name, training in (("static_small", False), ("tennis_dense", False), ("minecraft_small", False), ("tennis_dense", True)):
    os.environ["PE_BWD_TC"] = "0"; os.environ["PE_SAVE_FORWARD"] = "0"
    _, _, ref_in, ref_par = run_backward(name, training)
    os.environ["PE_BWD_TC"] = "1"; os.environ["PE_SAVE_FORWARD"] = "1"
    for wlo in ("1", "0"):
        os.environ["PE_BWD_CHAIN_WLO"] = wlo
        golden, loss, got_in, got_par = run_backward(name, training, precision="fp16x3")
        worst_g = max(compare_grads(got_in, got_par, golden, 0.0).values())
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
        worst_p = max(rel(got_par[k], ref_par[k]) for k in got_par if "nerf_model" in k or "ray_bender" in k)
        worst_i = max(rel(got_in[k], ref_in[k]) for k in got_in)
        print(json.dumps({"scene": name, "train": training, "wlo": wlo, "worst_vs_golden": round(worst_g, 4), "params_vs_fp32": round(worst_p, 4), "inputs_vs_fp32": round(worst_i, 4)}), flush=True)
for wlo in ("1", "0"):
    os.environ["PE_BWD_CHAIN_WLO"] = wlo
    for dense in (False, True):
        r = bench.train_step_report(torch.device("cuda", 0), dense)
        print(json.dumps({"wlo": wlo, "dense": dense, "fwd_bwd_ms": round(r["fwd_bwd_ms"], 3)}), flush=True)
