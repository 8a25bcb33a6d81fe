"""Node visits and triangle tests per ray of the C4 workload, split into the primary wave and the first reflections
(RTS_COUNT_NODES, which walks the BVH for the primary wave too; the difference between a maxRefl = 0 and a maxRefl = 1
trace of the same pulse)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rts_b200 import scenes, lib as L

def main():
    ms = scenes.terrain_scene(n=4096, n_rx=1)
    eng = L.Engine(0)
    eng.set_targets(ms.base)
    eng.set_poses(*ms.poses(1))
    out = {}
    for refl in (0, 1):
        sp = ms.spec_for(1)
        sp.max_refl = refl
        st = eng.trace(sp, L.RTS_OUT_BINS | L.RTS_COUNT_NODES | L.RTS_NO_REUSE)
        out[refl] = st
        print(json.dumps({"max_refl": refl, "segments": st["segments"], "nodes": st["nodes_visited"], "tris": st["tris_tested"],
                          "waves": [(round(a, 3), b) for a, b in eng.wave_profile()]}))
    a, b = out[0], out[1]
    seg = b["segments"] - a["segments"]
    print(json.dumps({"primary nodes/ray": round(a["nodes_visited"] / a["segments"], 2), "primary tris/ray": round(a["tris_tested"] / a["segments"], 2),
                      "later segments": seg, "later nodes/seg": round((b["nodes_visited"] - a["nodes_visited"]) / seg, 2),
                      "later tris/seg": round((b["tris_tested"] - a["tris_tested"]) / seg, 2), "depth": eng.bvh_info().__class__.__name__}))
    if len(sys.argv) > 1:
        return
    print("L2 read GB/s", [round(eng.probe_read_bandwidth(mb << 20, 100), 1) for mb in (16, 32, 48, 64, 96)], "HBM read GB/s", round(eng.probe_read_bandwidth(2 << 30, 3), 1))

if __name__ == "__main__":
    main()
