#!/bin/bash
# Tuning build of librts_b200.so with extra defines for trace.cu:
#   tools/build_variant.sh <tag> "<-D...>"   ->  rts_b200/variants/librts_b200_<tag>.so
# Select it at run time with RTS_B200_LIB=<path> (rts_b200/lib.py).  The .so files are git-ignored.
set -e
cd "$(dirname "$0")/../rts_b200/csrc"
mkdir -p ../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 $ARCH -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math"
make -s all
nvcc $FLAGS $2 -Xptxas -v -c trace.cu -o ../variants/trace_$1.o 2> ../variants/trace_$1.ptxas.log
nvcc $ARCH -shared -ccbin /usr/bin/g++ -o ../variants/librts_b200_$1.so api.o bvh.o ../variants/trace_$1.o aggregate.o comm.o host_mesh.o -cudart static -ldl
rm -f ../variants/trace_$1.o
grep -A2 "Function properties.*k_traverseILb0" ../variants/trace_$1.ptxas.log | grep "Used\|spill" | tr '\n' ' '; echo
