"""Throughput of the BASELINE.json configs C1-C3 at their full sizes on one GPU, and of one rank's share of a C5 pulse
(run under gpurun); C4 is bench.py."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rts_b200 import scenes, lib as L

ORACLE = "--oracle" in sys.argv          # also time the CPU oracle (BVH mode, OpenMP, all host cores) on the same pulse
if ORACLE:
    sys.argv.remove("--oracle")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O


def oracle_line(targets, spec):
    """the oracle on the GPU box's host cores, same pulse, in the same run."""
    if not ORACLE:
        return {}
    t0 = time.time()
    bins, st = O.trace_bins(targets, spec, use_bvh=True)
    return {"oracle_Mrays/s": round(st["primary_rays"] / st["ms_trace"] / 1e3, 3), "oracle_cores": int(O.oracle().orc_num_threads()),
            "oracle_segments": st["segments"], "oracle_bins": len(bins), "oracle_wall_s": round(time.time() - t0, 1)}

def main():
    eng = L.Engine(0)
    cases = [("C1 flat plate 256x256, 1 bounce", scenes.flat_plate(n=256)),
             ("C2 trihedral 1000x1000, 3 bounces", scenes.trihedral(n=1000)),
             ("C2 trihedral cubic 100^3 (reference launch shape)", scenes.trihedral(n=100, cubic=True)),
             ("C3 ship 100k triangles, refraction, 4096x4096, 4 Rx", scenes.ship(n=4096, hull_res=200))]
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, (t, s) in cases:
        if only and not name.startswith(only):
            continue
        eng.set_targets(t)
        best = None
        for rep in range(4):
            st = eng.trace(s, L.RTS_OUT_BINS | L.RTS_NO_REUSE)   # every repetition from scratch
            if best is None or st["ms_trace"] < best["ms_trace"]:
                best = st
        bins = eng.bins()
        line = {"config": name, "triangles": int(sum(len(x.tris) for x in t)), "rays": best["primary_rays"], "segments": best["segments"],
                "refracted": best["refracted"], "captured": best["captured"], "bins": len(bins), "ms_trace": round(best["ms_trace"], 3),
                "Mrays/s": round(best["primary_rays"] / best["ms_trace"] / 1e3, 1),
                "Msegments/s": round(best["segments"] / best["ms_trace"] / 1e3, 1),
                "waves": [(round(a, 3), b) for a, b in eng.wave_profile()]}
        line.update(oracle_line(t, s))
        if "oracle_Mrays/s" in line:
            line["gpu_over_oracle"] = round(line["Mrays/s"] / line["oracle_Mrays/s"], 1)
        print(json.dumps(line), flush=True)
    if not only or only == "C5":
        c5_shard(eng)

def c5_shard(eng, world=8, rank=3, pulses=4):
    """C5: 8 receivers, (1,10000,10000) = 1e8 rays per pulse sharded round-robin over 8 GPUs; this is rank 3's share of
    consecutive pulses (12.5M rays each, moving targets, refit), from scratch, as bench.py steps."""
    ms = scenes.terrain_scene(n=10000, n_rx=8, nz=10000)
    eng.set_targets(ms.base)
    best, segs = None, 0
    for p in range(pulses):
        eng.set_poses(*ms.poses(p))
        sp = ms.spec_for(p)
        sp.ray_begin, sp.ray_count, sp.ray_stride = rank, 0, world
        st = eng.trace(sp, L.RTS_OUT_BINS | L.RTS_NO_REUSE)
        if best is None or st["ms_trace"] < best["ms_trace"]:
            best = st
    bins = eng.bins()
    line = {"config": f"C5 multistatic: rank {rank} of {world}'s share of a (1,10000,10000) pulse, 8 Rx, 1M-triangle terrain + movers",
            "triangles": int(sum(len(x.tris) for x in ms.base)), "rays": best["primary_rays"], "segments": best["segments"],
            "captured": best["captured"], "bins": len(bins), "ms_trace": round(best["ms_trace"], 3),
            "Mrays/s": round(best["primary_rays"] / best["ms_trace"] / 1e3, 1),
            "Msegments/s": round(best["segments"] / best["ms_trace"] / 1e3, 1),
            "waves": [(round(a, 3), b) for a, b in eng.wave_profile()]}
    if ORACLE:      # every 4th ray of the share keeps the CPU leg to a few seconds
        sp = ms.spec_for(pulses - 1)
        sp.ray_begin, sp.ray_count, sp.ray_stride = rank, 0, world * 4
        line.update(oracle_line(ms.world_targets(pulses - 1), sp))
        line["oracle_sample"] = "every 4th ray of the share"
        line["gpu_over_oracle"] = round(line["Mrays/s"] / line["oracle_Mrays/s"], 1)
    print(json.dumps(line), flush=True)

if __name__ == "__main__":
    main()
