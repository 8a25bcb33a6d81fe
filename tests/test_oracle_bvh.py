"""The oracle's own BVH (used only so that it finishes on 100k-1M triangle scenes) against its
exhaustive search — the definitional closest hit — on sub-sampled rays."""
import numpy as np

import oracle_api as O
from rts_b200 import scenes


def _same_rays(a, b, spec, idx):
    M, R3 = spec.slots, spec.rays
    rows = np.concatenate([idx + k * R3 for k in range(M)])
    for f in a["results"].dtype.names:
        assert np.ascontiguousarray(a["results"][f][rows]).tobytes() == np.ascontiguousarray(b["results"][f][rows]).tobytes(), f
    assert np.array_equal(a["tri_path"][rows], b["tri_path"][rows])
    assert np.array_equal(a["targ_intersect"][rows], b["targ_intersect"][rows])


def test_bvh_equals_brute_force_ship_subsample():
    targets, spec = scenes.ship(n=96, hull_res=24)
    spec.ray_stride = 7
    a = O.trace(targets, spec, use_bvh=False)
    b = O.trace(targets, spec, use_bvh=True)
    idx = np.arange(0, spec.rays, 7)
    _same_rays(a, b, spec, idx)
    assert a["stats"]["segments"] == b["stats"]["segments"] and a["stats"]["hits"] == b["stats"]["hits"] > 0
    assert a["stats"]["refracted"] > 0


def test_bvh_equals_brute_force_terrain_subsample():
    ms = scenes.terrain_scene(n=64, cells_x=60, cells_y=30, movers=4)
    targets, spec = ms.world_targets(3), ms.spec_for(3)
    a = O.trace(targets, spec, use_bvh=False)
    b = O.trace(targets, spec, use_bvh=True)
    _same_rays(a, b, spec, np.arange(spec.rays))
    assert a["stats"]["hits"] > 1000


def test_sharded_trace_covers_the_same_rays():
    """ray_begin / ray_count / ray_stride select primary rays without changing any ray's result."""
    targets, spec = scenes.trihedral(n=40)
    full = O.trace(targets, spec)
    spec.ray_begin, spec.ray_count = 500, 700
    part = O.trace(targets, spec)
    sel = slice(500, 1200)
    for f in full["results"].dtype.names:
        assert np.array_equal(full["results"][f][sel], part["results"][f][sel])
    assert part["stats"]["primary_rays"] == 700
    assert (part["results"]["reflDepth"][:500] == 0).all() and (part["tri_path"][:500] == -1).all()


def test_bvh_keeps_rays_that_run_along_a_shared_triangle_edge():
    """An odd launch grid puts one column of rays in the plane y = 0 (d.y ~ 1e-17), which is a cell boundary of the
    terrain: the fp64 triangle test accepts hits there by a rounding error's margin, and node boxes that end exactly on
    y = 0 used to prune them (found by tests/test_gpu_fuzz.py, where the GPU agreed with exhaustive search)."""
    ms = scenes.terrain_scene(n=115, cells_x=47, cells_y=24, movers=3, n_rx=2, seed=0x52545301 + 51)
    targets, spec = ms.world_targets(0), ms.spec_for(0)
    a = O.trace(targets, spec, use_bvh=False)
    b = O.trace(targets, spec, use_bvh=True)
    _same_rays(a, b, spec, np.arange(spec.rays))
    for k in ("segments", "hits", "shaded_hits", "captured"):
        assert a["stats"][k] == b["stats"][k], k
