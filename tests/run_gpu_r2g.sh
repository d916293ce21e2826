#!/bin/bash
# ncu --set full of the compositor kernels (sparse Tennis train step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:pe_composite -s 2 -c 2 -o gpurun_out/r2g_composite -f python tests/gpu_train_step.py sparse 1 > gpurun_out/r2g_ncu.log 2>&1
tail -3 gpurun_out/r2g_ncu.log
