#!/bin/bash
# The measurement pass of a round, run on the GPU box (gpurun -- 'bash tools/profile_round.sh [tag]'):
#   1. the bench line (no profiler; its roofline counters come from the ncu child bench.py spawns itself)  -> gpurun_out/bench_n1.json
#   2. ncu launch list of a short bench run                                                  -> gpurun_out/launches.csv
#   3. one ncu --set full capture of every wave / projection kernel of one from-scratch step  -> gpurun_out/prof.ncu-rep
#   4. the other BASELINE.json configs at full size                                          -> gpurun_out/config_probe.jsonl
# Then, in the dev container: python tools/ncu_summary.py rNN; python tools/ncu_hotspots.py ... (see profiles/).
# Numbers printed by the runs under ncu are never bench values.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_err.log || { tail -20 gpurun_out/bench_err.log; exit 1; }
tail -c 1500 gpurun_out/bench_n1.json
python bench.py --steps 2 --warmup 1 --quick > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --quick > gpurun_out/ncu1.log 2>&1
# one whole step's wave / projection kernels (8 launches per from-scratch step: directions, 2 footprint kernels, k_primary_follow,
# the BVH primary wave that returns at once, 3 later waves): skip the warm-up step and the first timed step of the value leg
ncu --set full --clock-control none --import-source on -k 'regex:k_wave|k_raster|k_primary|k_traverse|k_shade' -s 16 -c 8 -f -o gpurun_out/prof -- \
    python bench.py --steps 2 --warmup 1 --quick > gpurun_out/ncu2.log 2>&1
python tools/config_probe.py > gpurun_out/config_probe.jsonl 2> gpurun_out/config_probe.err
cut -c1-100 gpurun_out/config_probe.jsonl
