mkdir -p gpurun_out
python tools/shard_probe.py 1 2>&1 | cut -c1-140
python tools/config_probe.py | cut -c1-330
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
. tools/ab.sh
run coop2
