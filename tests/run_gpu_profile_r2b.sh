#!/bin/bash
# Evidence for profiles/r2_aware_rounding.md: mask sweeps with / without the activation-aware weight stream, ncu metrics and full-set
# capture of the headline instantiation in the shipped mixed mode, launch list of the bench step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python tests/gpu_mixed_sweep.py 0x000 0x080 0x0C0 0x0E0 0x0F8 > gpurun_out/r2x_sweep_aware.jsonl 2> gpurun_out/r2x_sweep_aware.err
echo "sweep aware exit $?"
PE_TC_AWARE=0 timeout 300 python tests/gpu_mixed_sweep.py 0x000 0x0C0 0x0F8 > gpurun_out/r2x_sweep_zero_sum.jsonl 2> gpurun_out/r2x_sweep_zero_sum.err
echo "sweep zero-sum exit $?"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:pe_field_tc_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2x_tc_fullsize_mixed_aware.csv python tests/profile_tc.py 256 mixed 2 > /dev/null 2>&1
echo "metrics exit $?"; grep -v "^==" gpurun_out/r2x_tc_fullsize_mixed_aware.csv | awk -F'","' 'NR>1{print $(NF-2), $(NF-1), $NF}'
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pe_field_tc_kernel -s 1 -c 1 -o gpurun_out/r2x_mixed_aware_full -f python tests/profile_tc.py 256 mixed 2 > gpurun_out/r2x_ncu_full.log 2>&1
echo "full set exit $?"; ls -la gpurun_out/r2x_mixed_aware_full.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2x_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --quick > gpurun_out/r2x_launch_bench.log 2>&1
echo "launch list exit $?"
