#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/train_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/train_pytest.log
timeout 300 python - <<'PY'
import sys, torch
sys.path[:0]=['.','tests','tests/golden']
import bench
torch.cuda.set_device(0)
for dense in (False, True):
    print(bench.train_step_report(torch.device('cuda',0), dense))
PY
