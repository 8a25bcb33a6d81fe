mkdir -p gpurun_out
python -m pytest tests/test_gpu_variants.py -x -q > gpurun_out/pytest_var.log 2>&1; echo "variants rc=$?"; tail -12 gpurun_out/pytest_var.log
. tools/ab.sh
run noq
