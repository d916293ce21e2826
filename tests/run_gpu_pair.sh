#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== single-CTA kernel (refactored)"
PE_TC_KERNEL=1 timeout 200 python tests/gpu_diag.py fp16/static fp16x2/static 2>&1 | grep -E "^(ok|FAIL)"
echo "== CTA-pair kernel"
PE_TC_KERNEL=2 timeout 120 python tests/gpu_diag.py fp16/static fp16x2/static tc_vs_fp32 perturb_fp16 2>&1 | grep -E "^(ok|FAIL)|rror" | cut -c1-400
echo "exit $?"
for p in fp16 fp16x2; do
  PE_TC_KERNEL=2 timeout 120 python bench.py --precision $p --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_$p.json 2> gpurun_out/bench2_$p.err
  echo "== bench pair $p exit $?"; python -c "
import json
d=json.load(open('gpurun_out/bench2_$p.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks')}, d['roofline']['frac'], d['e2e']['value'])"; tail -2 gpurun_out/bench2_$p.err
done
