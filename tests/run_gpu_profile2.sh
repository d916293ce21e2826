#!/bin/bash
# full-set capture of the field kernel on a quarter frame (128x128 rays x 128 samples; same per-tile work, 4x fewer tiles)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/quick_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/quick_pytest.log
for p in fp16 fp16x2; do
timeout 240 python bench.py --precision $p --steps 20 --warmup 3 --no-cpu-baseline --quick > gpurun_out/quick_$p.json 2> gpurun_out/quick_$p.err
python -c "
import json; d=json.load(open('gpurun_out/quick_$p.json')); print('$p', d['ms_per_step'], d['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
done
for p in fp16; do
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pe_field_tc_kernel -s 1 -c 1 -o gpurun_out/r1_tc_fold_$p -f python tests/profile_tc.py 128 $p 2 > gpurun_out/ncu_$p.log 2>&1
echo "full $p exit $?"; tail -3 gpurun_out/ncu_$p.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
