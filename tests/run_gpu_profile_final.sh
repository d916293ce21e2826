#!/bin/bash
# Evidence for profiles/: launch list of the bench step, per-kernel metrics on the full frame, full-set capture on a quarter frame
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --quick > gpurun_out/r1c_launch_bench.log 2>&1
echo "launch list exit $?"
for p in fp16 fp16x2; do
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum --clock-control none -k regex:pe_field_tc_kernel -s 1 -c 1 --csv --log-file gpurun_out/r1c_tc_fullsize_$p.csv python tests/profile_tc.py 256 $p 2 > /dev/null 2>&1
echo "metrics $p exit $?"; grep -v "^==" gpurun_out/r1c_tc_fullsize_$p.csv | awk -F'","' 'NR>1{print $(NF-2), $(NF-1), $NF}'
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pe_field_tc_kernel -s 1 -c 1 -o gpurun_out/r1c_tc_quarter_fp16 -f python tests/profile_tc.py 128 fp16 2 > gpurun_out/ncu_r1c.log 2>&1
echo "full set exit $?"; ls -la gpurun_out/r1c_tc_quarter_fp16.ncu-rep
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1c_launches_eval_dense.csv python tests/profile_eval.py 1 fp16 > /dev/null 2>&1
echo "eval launch list exit $?"
