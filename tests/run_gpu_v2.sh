#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for group in "umma_gemm" "fp16/ perturb_fp16 tc_vs_fp32" "fp16x2"; do
  tag=$(echo $group | tr ' /' '__')
  timeout 300 python tests/gpu_diag.py $group > gpurun_out/diag_$tag.log 2>&1
  echo "== $group exit $?"
  sed -n '/==== SUMMARY ====/,$p' gpurun_out/diag_$tag.log
done
for p in fp16 fp16x2; do
  timeout 300 python bench.py --precision $p --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err
  echo "== bench $p exit $?"; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$p.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'], d['e2e']['value'])"; tail -2 gpurun_out/bench_$p.err
done
