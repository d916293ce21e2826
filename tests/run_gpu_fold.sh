#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/fold_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/fold_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -14
for f in 1 0; do
for p in fp16 fp16x2 fp16x3; do
  PE_TC_FOLD=$f timeout 300 python bench.py --precision $p --steps 20 --warmup 3 --no-cpu-baseline --quick > gpurun_out/fold${f}_$p.json 2> gpurun_out/fold${f}_$p.err
  echo "== fold=$f $p exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/fold${f}_$p.json")); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["gpu_launches"])
except Exception as e: print("ERR", e)
PY
  tail -3 gpurun_out/fold${f}_$p.err
done; done
