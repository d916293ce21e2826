#!/bin/bash
# Builds libpe_b200.so for sm_100a in-tree (the .so travels to the GPU box with the repo snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $PE_EXTRA_FLAGS"
mkdir -p build
pids=()
for f in pe_abi pe_misc pe_composite pe_field_fp32 pe_field_tc pe_field_bwd pe_bwd_tc pe_backward; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer build/$f.o)" ] || [ ../../include/pe_b200.h -nt build/$f.o ]; then
    ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libpe_b200.so build/pe_abi.o build/pe_misc.o build/pe_composite.o build/pe_field_fp32.o build/pe_field_tc.o build/pe_field_bwd.o build/pe_bwd_tc.o build/pe_backward.o -lcudart
echo "built $(realpath ../libpe_b200.so)"
