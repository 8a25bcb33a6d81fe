"""Quick GPU-vs-oracle check on small scenes (run under gpurun). Prints a JSON summary."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_api as O
import parity
from rts_b200 import scenes, lib as L

def main():
    eng = L.Engine(0)
    out = []
    cases = [
        ("plate16c", *scenes.flat_plate(n=16, cubic=True)),
        ("plate256", *scenes.flat_plate(n=256)),
        ("trihedral200", *scenes.trihedral(n=200)),
        ("slab64", *scenes.slab(n=64)),
        ("slab64-interp", *scenes.slab(n=64, interpolate=True, refr_index=1.3, max_refl=3)),
    ]
    for name, t, s in cases:
        t0 = time.time()
        orc = O.trace(t, s)
        t1 = time.time()
        recs, gbins, st = parity.run_gpu_records(eng, t, s)
        t2 = time.time()
        c = parity.compare_records(recs, orc, s, name)
        obins, ost = O.trace_bins(t, s, use_bvh=False)
        cb = parity.compare_bins(gbins, obins)
        c.update({"bins": cb, "gpu_stats": {k: st[k] for k in ("segments", "hits", "shaded_hits", "captured", "refracted", "nodes_visited", "tris_tested", "ms_trace")},
                  "orc_stats": {k: orc["stats"][k] for k in ("segments", "hits", "shaded_hits", "captured", "refracted")},
                  "t_oracle": t1 - t0, "t_gpu": t2 - t1})
        print(json.dumps(c))
        out.append(c)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
