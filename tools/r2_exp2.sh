#!/bin/bash
# round 2, experiment 2: quantised 32-byte nodes (fused and split), parity + A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
. tools/ab.sh
run fusedQ RTS_NO_SPLIT=1
run splitQ
run fused_noq RTS_NO_SPLIT=1 RTS_B200_LIB=rts_b200/variants/librts_b200_noq.so
run split_noq RTS_B200_LIB=rts_b200/variants/librts_b200_noq.so
for v in w7 w8; do run fusedQ_$v RTS_NO_SPLIT=1 RTS_B200_LIB=rts_b200/variants/librts_b200_$v.so; done
