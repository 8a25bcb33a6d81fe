"""A small tour of every kernel of the library, for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rts_b200 import scenes, lib as L

def main():
    eng = L.Engine(0)
    for name, (t, s) in {"plate": scenes.flat_plate(n=48), "trihedral": scenes.trihedral(n=64), "slab": scenes.slab(n=48),
                         "slab-smooth": scenes.slab(n=32, interpolate=True, refr_index=1.3, max_refl=3)}.items():
        eng.set_targets(t)
        st = eng.trace(s, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS | L.RTS_COUNT_NODES)
        res, ti, rcs, tp = eng.records()
        bins, resp = eng.bins(), eng.responses()
        rr = eng.received()
        if len(rr[0]):
            eng.aggregate(rr[0], rr[1], s.cspeed, s.carrier, ray_total=s.ray_total)
        eng.set_option("no_raster", 1)
        eng.trace(s, L.RTS_OUT_BINS)
        eng.set_option("no_raster", 0)
        print(name, st["segments"], st["hits"], len(bins), len(resp), len(rr[0]))
    ms = scenes.terrain_scene(n=96, cells_x=64, cells_y=40, movers=6, n_rx=2)     # 5120 + mover triangles: partial refit path
    eng.set_targets(ms.base)
    for p in range(4):
        eng.set_poses(*ms.poses(p))
        eng.trace(ms.spec_for(p), L.RTS_OUT_BINS | L.RTS_ASYNC)
    print("terrain", eng.stats()["segments"], len(eng.bins()), eng.check_bvh(), eng.bvh_info().builds)
    eng.rebuild()
    sp = ms.spec_for(3)
    sp.ray_begin, sp.ray_stride = 1, 3
    print("shard", eng.trace(sp, L.RTS_OUT_BINS)["segments"])
    eng.close()

if __name__ == "__main__":
    main()
