#!/bin/bash
# multi-GPU checks of a round (gpurun --gpus N -- 'bash tools/r2_multi.sh N'): peer-exchange tests, then bench.py at N with
# both exchanges, weak scaling; strong and pulse sharding with the peer exchange
N=${1:-2}
mkdir -p gpurun_out
[ "${2:-}" = "notest" ] || timeout 900 python -m pytest tests/test_gpu_peer_exchange.py tests/test_native_host.py -x -q -s > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; grep 'exchange device' gpurun_out/pytest_multi.log; tail -5 gpurun_out/pytest_multi.log
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N "$@"; }
for ex in peer nccl; do
  tr --steps 64 --warmup 8 --quick --exchange $ex > gpurun_out/bench_n${N}_$ex.json 2> gpurun_out/bench_n${N}_$ex.err || tail -5 gpurun_out/bench_n${N}_$ex.err
  python -c "
import json; t=open('gpurun_out/bench_n${N}_$ex.json').read(); d=json.loads(t[t.index('{\"metric'):].splitlines()[0]); print('$ex', 'N', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['bin_exchange'])"
done
for mode in strong pulse; do
  tr --steps 64 --warmup 8 --quick --scaling $mode > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err || tail -5 gpurun_out/bench_n${N}_$mode.err
  python -c "
import json; t=open('gpurun_out/bench_n${N}_$mode.json').read(); d=json.loads(t[t.index('{\"metric'):].splitlines()[0]); print('$mode', 'N', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['bin_exchange'])"
done
