#!/bin/bash
# full GPU suite + eval-frame / train-step timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2e_pytest.log
timeout 300 python tests/gpu_eval_frames.py > gpurun_out/r2e_eval.jsonl 2>&1; cat gpurun_out/r2e_eval.jsonl
timeout 300 python tests/gpu_train_step.py both > gpurun_out/r2e_train.jsonl 2>&1; cut -c1-400 gpurun_out/r2e_train.jsonl | grep -o '"fwd_ms.*fwd_bwd_ms": [0-9.]*'
