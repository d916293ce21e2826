#!/bin/bash
# eval-frame timings + launch list of the T-frame / Tennis eval frames
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tests/gpu_eval_frames.py > gpurun_out/r2d_eval.jsonl 2>&1; cat gpurun_out/r2d_eval.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2d_launches_eval.csv python tests/gpu_eval_frames.py > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
