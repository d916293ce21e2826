#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/launch_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pe_field_tc_kernel -s 1 -c 1 -o gpurun_out/prof_tc_r1b -f python tests/profile_tc.py 256 fp16 2 > gpurun_out/ncu_b.log 2>&1
echo "full fp16 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:pe_field_tc_kernel -s 1 -c 1 -o gpurun_out/prof_tc_r1c -f python tests/profile_tc.py 256 fp16x2 2 > gpurun_out/ncu_c.log 2>&1
echo "full fp16x2 exit $?"
ls -la gpurun_out/*.ncu-rep
