#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in fp16 fp16x2; do
  timeout 300 python bench.py --precision $p --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err
  echo "== bench $p exit $?"; cat gpurun_out/bench_$p.json; tail -3 gpurun_out/bench_$p.err
done
timeout 300 python bench.py --precision fp32 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
echo "== bench fp32 exit $?"; cat gpurun_out/bench_fp32.json; tail -3 gpurun_out/bench_fp32.err
timeout 300 python tests/gpu_diag.py perturb_fp16 > gpurun_out/diag_perturb.log 2>&1; sed -n '/==== SUMMARY/,$p' gpurun_out/diag_perturb.log
