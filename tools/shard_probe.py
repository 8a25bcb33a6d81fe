"""Per-rank cost of a ray-sharded C4 pulse on ONE GPU: the shares of rank 0 .. N-1 of a (1, 4096*N, 4096) launch traced one after
the other by the same engine (is a share slower because of its lattice offset, or because of the GPU it runs on?)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rts_b200 import scenes, lib as L

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ms = scenes.terrain_scene(n=4096 * world, n_rx=1, nz=4096)
eng = L.Engine(0)
eng.set_targets(ms.base)
order = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else range(world)
for rank in order:
    best = None
    for p in range(6):
        eng.set_poses(*ms.poses(p))
        sp = ms.spec_for(p)
        sp.ray_begin, sp.ray_count, sp.ray_stride = rank, 0, world
        st = eng.trace(sp, L.RTS_OUT_BINS | L.RTS_NO_REUSE)
        if p >= 2 and (best is None or st["ms_trace"] < best[0]):
            best = (st["ms_trace"], [round(a, 4) for a, b in eng.wave_profile()], eng.follow_profile(), st["segments"], st["hits"])
    print(json.dumps({"rank": rank, "world": world, "ms_trace": round(best[0], 4), "waves": best[1], "follow_ms": round(best[2], 4), "segments": best[3], "hits": best[4]}), flush=True)
