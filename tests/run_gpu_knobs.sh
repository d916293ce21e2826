#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do
  PE_TC_DEBUG=$d timeout 200 python bench.py --precision fp16 --steps 10 --warmup 2 --no-cpu-baseline > gpurun_out/knob_$d.json 2> gpurun_out/knob_$d.err
  python -c "
import json
d=json.load(open('gpurun_out/knob_$d.json')); print('dbg=$d ms_per_step', round(d['ms_per_step'],3))"
done
