"""When do the direction pass, the footprint kernels and the shading pass of consecutive from-scratch pulses of the bench
scene run?  (knob debug_timeline; printed to stderr by rts_sync).  gpurun -- python tools/timeline_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rts_b200 import lib as L, scenes

ms = scenes.terrain_scene(n=4096, n_rx=1, nz=4096)
eng = L.Engine(0)
eng.set_targets(ms.base)
flags = L.RTS_OUT_BINS | L.RTS_ASYNC | L.RTS_NO_REUSE
prepared = [(L.Engine.pack_poses(*ms.poses(p)), eng.prepare(ms.spec_for(p))) for p in range(12)]
for rep in range(2):
    if rep == 1:
        eng.set_option("debug_timeline", 1)
    for (poses, n), cp in prepared[:10]:
        eng.set_poses_packed(poses, n)
        eng.trace_prepared(cp, flags)
    eng.sync()
