#!/bin/bash
# one ncu --set full capture of the fused primary-shade + first-reflection kernel (third launch of it: steady state)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_primary_follow -s 2 -c 1 -f -o gpurun_out/prof_follow -- \
    python bench.py --steps 2 --warmup 1 --quick > gpurun_out/ncu_follow.log 2>&1
ls -la gpurun_out/*.ncu-rep
