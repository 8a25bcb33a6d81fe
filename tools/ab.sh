# A/B helper for gpurun sessions: `. tools/ab.sh; run <tag> [VAR=value ...]` prints value, ms/step, per-wave times and the
# second wave's two kernels of a short bench run (value and e2e legs only)
run() { tag=$1; shift; env "$@" python bench.py --steps 32 --warmup 4 --quick 2>gpurun_out/ab_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$tag', 'value', d['value'], 'ms/step', d['ms_per_step'], 'waves', r['waves_ms_per_step'], 'follow', r.get('k_primary_follow_ms'), 'split', r['second_wave_kernels_ms'], 'e2e', d['e2e']['value'], flush=True)" || tail -3 gpurun_out/ab_$tag.err; }
