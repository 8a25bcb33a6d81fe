"""Per-wave timing and traversal counters of the C4 workload (run under gpurun)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rts_b200 import scenes, lib as L

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ms = scenes.terrain_scene(n=n, n_rx=1)
    eng = L.Engine(0)
    eng.set_targets(ms.base)
    info = eng.bvh_info()
    print(f"build {info.ms_build:.2f} ms nodes {info.n_nodes}")
    for p in range(reps):
        eng.set_poses(*ms.poses(p))
        st = eng.trace(ms.spec_for(p), L.RTS_OUT_BINS)
        wp = eng.wave_profile()
        bi = eng.bvh_info()
        print(json.dumps({"pulse": p, "builds": bi.builds, "sah": round(bi.sah_cost / max(bi.sah_at_build, 1e-300), 4), "refit_ms": round(eng.bvh_info().ms_refit, 3), "trace_ms": round(st["ms_trace"], 3),
                          "Mrays/s": round(st["primary_rays"] / st["ms_trace"] / 1e3, 1),
                          "seg/ray": round(st["segments"] / st["primary_rays"], 3),
                          "nodes/seg": round(st["nodes_visited"] / st["segments"], 1), "tris/seg": round(st["tris_tested"] / st["segments"], 2),
                          "waves": [(round(a, 2), b) for a, b in wp], "captured": st["captured"]}))
    eng.rebuild()
    st = eng.trace(ms.spec_for(reps - 1), L.RTS_OUT_BINS)
    print("after rebuild: trace_ms", round(st["ms_trace"], 3), "nodes/seg", round(st["nodes_visited"] / st["segments"], 1))

if __name__ == "__main__":
    main()
