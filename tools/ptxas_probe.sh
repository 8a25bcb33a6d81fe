#!/bin/bash
# Registers / spills of one kernel of trace.cu under extra defines (no GPU needed):
#   tools/ptxas_probe.sh "-DRTS_TRAV_MIN_BLOCKS=10" k_traverseILb0
cd "$(dirname "$0")/../rts_b200/csrc"
mkdir -p ../../build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -ccbin /usr/bin/g++ \
     -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -Xptxas -v $1 -c trace.cu -o ../../build/trace_probe.o 2>&1 |
  grep -A2 "Function properties for.*$2" | grep -v "^--" | sed 's/_ZN40_GLOBAL__N__[0-9a-f_]*trace_cu_[0-9a-f]*//'
