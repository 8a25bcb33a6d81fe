#!/bin/bash
# Tuning builds of librts_b200.so with different launch bounds for the wave kernel:
#   tools/build_variants.sh 68 67 66   ->  rts_b200/variants/librts_b200_<XY>.so   (X = later waves, Y = primary wave)
# Select one at run time with RTS_B200_LIB=<path> (rts_b200/lib.py).  The .so files are git-ignored.
set -e
cd "$(dirname "$0")/../rts_b200/csrc"
mkdir -p ../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 $ARCH -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math"
make -s all
for v in "$@"; do
  x=${v:0:1}; y=${v:1:1}; extra=${v:2}
  nvcc $FLAGS -DRTS_WAVE_MIN_BLOCKS=$x -DRTS_WAVE_MIN_BLOCKS_PRIMARY=$y $EXTRA_DEFS -c trace.cu -o ../variants/trace_$v.o &
done
wait
for v in "$@"; do
  nvcc $ARCH -shared -ccbin /usr/bin/g++ -o ../variants/librts_b200_$v.so api.o bvh.o ../variants/trace_$v.o aggregate.o comm.o host_mesh.o -cudart static -ldl
  rm -f ../variants/trace_$v.o
done
ls -la ../variants
