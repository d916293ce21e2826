#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PE_TC_TIMELINE=1 timeout 300 python tests/profile_tc.py 256 ${1:-fp16} 2 2>&1 | grep "PE_TC" | tail -11
