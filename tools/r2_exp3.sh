#!/bin/bash
# round 2, experiment 3: primary shading pass follows the first reflection in place (follow.cuh)
mkdir -p gpurun_out
. tools/ab.sh
run follow
run nofollow RTS_NO_FOLLOW=1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
