#!/bin/bash
# Runs the GPU diagnostics in separate processes (a hang in one kernel family must not hide the others).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for group in "umma_gemm" "fp32 perturb_fp32 train_fp32 ops" "fp16/ perturb_fp16 tc_vs_fp32" "fp16x2"; do
  tag=$(echo $group | tr ' /' '__')
  timeout 420 python tests/gpu_diag.py $group > gpurun_out/diag_$tag.log 2>&1
  echo "== $group exit $?"
  sed -n '/==== SUMMARY ====/,$p' gpurun_out/diag_$tag.log
  grep -m3 -i "error\|Traceback" gpurun_out/diag_$tag.log | head -5
done
