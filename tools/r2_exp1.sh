#!/bin/bash
# round 2, experiment 1: parity of the split waves + A/B of k_traverse launch bounds and refill thresholds
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
. tools/ab.sh
run fused RTS_NO_SPLIT=1
run split_t8f8
for v in t8f16 t8f32 t8f4 t7f8 t9f8 t9f16 t10f8; do
  run $v RTS_B200_LIB=rts_b200/variants/librts_b200_$v.so
done
