#!/bin/bash
# full GPU suite + train-step timing + launch list of the sparse / dense train step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2c_pytest.log
timeout 300 python tests/gpu_train_step.py both > gpurun_out/r2c_train.jsonl 2>&1; cat gpurun_out/r2c_train.jsonl
for w in sparse dense; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c_launches_$w.csv python tests/gpu_train_step.py $w 1 > gpurun_out/r2c_ncu_$w.log 2>&1
done
