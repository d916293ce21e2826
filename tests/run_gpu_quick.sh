#!/bin/bash
# quick kernel iteration: parity tests + fp16/fp16x2 bench + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/quick_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/quick_pytest.log
for p in fp16 fp16x2; do
  timeout 300 python bench.py --precision $p --steps 20 --warmup 3 --no-cpu-baseline --quick > gpurun_out/quick_$p.json 2> gpurun_out/quick_$p.err
  echo "== $p exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/quick_$p.json")); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
except Exception as e: print("ERR", e)
PY
  tail -3 gpurun_out/quick_$p.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/quick_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --quick > gpurun_out/quick_launch_bench.log 2>&1
grep -v "^==" gpurun_out/quick_launches.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[1:]:
    if len(r)>vi: print(r[ki][:60], r[vi], r[ui])
" | tail -16
