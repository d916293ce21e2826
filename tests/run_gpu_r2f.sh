#!/bin/bash
# compositor check: composite-related tests + eval/train timings + launch list of the sparse / dense train step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python tests/gpu_eval_frames.py > gpurun_out/r2f_eval.jsonl 2>&1; cat gpurun_out/r2f_eval.jsonl
timeout 300 python tests/gpu_train_step.py both > gpurun_out/r2f_train.jsonl 2>&1; grep -o '"fwd_ms.*fwd_bwd_ms": [0-9.]*' gpurun_out/r2f_train.jsonl
for w in sparse dense; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:composite -c 40 --csv --log-file gpurun_out/r2f_launches_$w.csv python tests/gpu_train_step.py $w 1 > gpurun_out/r2f_ncu_$w.log 2>&1
grep composite gpurun_out/r2f_launches_$w.csv | tail -4 | awk -F'","' '{print $5, $NF}'
done
