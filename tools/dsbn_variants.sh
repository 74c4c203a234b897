#!/bin/bash
# Development tool: rebuild dsbn.cu with different (vectors per item, resident blocks) and time the kernels.
cd "$(dirname "$0")/../fpl-plus_b200" || exit 1
for cfg in "2 4" "4 3" "4 2" "1 4"; do
  set -- $cfg
  nvcc -DFPL_BWD_V=$1 -DFPL_BWD_BLOCKS=$2 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -c csrc/dsbn.cu -o build/dsbn.o || exit 1
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libfplplus_b200.so build/*.o || exit 1
  echo "== V=$1 BLOCKS=$2"
  REPS=20 python ../tools/dsbn_probe.py 2>&1 | head -5
done
