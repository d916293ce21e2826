#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:pe_field_tc_kernel -s 1 -c 1 --csv --log-file gpurun_out/r1_tc_fold_fullsize.csv python tests/profile_tc.py 256 fp16 2 > gpurun_out/ncu_fullsize.log 2>&1
echo "fullsize metrics exit $?"; grep -v "^==" gpurun_out/r1_tc_fold_fullsize.csv | cut -d, -f 5,13- | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:pe_field_tc_kernel -s 1 -c 1 --csv --log-file gpurun_out/r1_tc_fold_fullsize_x2.csv python tests/profile_tc.py 256 fp16x2 2 > gpurun_out/ncu_fullsize_x2.log 2>&1
echo "fullsize x2 metrics exit $?"; grep -v "^==" gpurun_out/r1_tc_fold_fullsize_x2.csv | cut -d, -f 5,13- | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --quick > gpurun_out/r1_launch_bench.log 2>&1
echo "launch list exit $?"
