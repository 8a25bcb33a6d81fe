mkdir -p gpurun_out
python -m pytest tests/test_gpu_tables.py -x -q > gpurun_out/pytest_tables.log 2>&1; echo "tables rc=$?"; tail -25 gpurun_out/pytest_tables.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
. tools/ab.sh
run final
python tools/config_probe.py C3 | cut -c1-330
