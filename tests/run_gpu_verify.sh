#!/bin/bash
# Full round-end style check: GPU tests, smoke, default bench line, reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/verify_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/verify_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/verify_smoke.log
timeout 600 python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; echo "bench exit $?"; cat gpurun_out/verify_bench.json; tail -3 gpurun_out/verify_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/verify_bench_ref.json 2> gpurun_out/verify_bench_ref.err; echo "ref exit $?"; cat gpurun_out/verify_bench_ref.json; tail -3 gpurun_out/verify_bench_ref.err
