"""Parity of the CUDA path (through the C-ABI, librts_b200.so) against the CPU oracle.

Bars (BASELINE.json acceptance): hit-triangle ids, bounce counts, path rows and every fp64 state
field bit-exact; rays the oracle flags as receiver-window edge cases (fp32 atan2f differs between
libdevice and glibc by a few ulp) are excluded from capture-dependent fields and counted; per-bin
counts exact, per-bin fp64 sums within 1e-5 relative (measured ~1e-15)."""
import glob
import os

import numpy as np
import pytest

import oracle_api as O
import parity
from rts_b200 import lib as L
from rts_b200 import scenes
from test_oracle_reference import CASES, GOLDEN

pytestmark = pytest.mark.gpu

SMALL = {
    "C1_plate_256": lambda: scenes.flat_plate(n=256),
    "plate_cubic_12": lambda: scenes.flat_plate(n=12, cubic=True),
    "plate_refl3": lambda: scenes.flat_plate(n=64, max_refl=3),
    "trihedral_300": lambda: scenes.trihedral(n=300),
    "trihedral_refl1": lambda: scenes.trihedral(n=100, max_refl=1),
    "trihedral_refl0": lambda: scenes.trihedral(n=50, max_refl=0),
    "slab": lambda: scenes.slab(n=96),
    "slab_interp": lambda: scenes.slab(n=64, interpolate=True, refr_index=1.3, max_refl=3),
    "slab_thin": lambda: scenes.slab(n=64, thickness=0.004, max_refl=1),
    "ship_small": lambda: scenes.ship(n=96, hull_res=32),
    "spheres_interp": lambda: scenes.spheres(n=96),
    "spheres_flat": lambda: scenes.spheres(n=64, interpolate=False, max_refl=3, max_refr=0),
}


@pytest.mark.parametrize("name", sorted(SMALL))
def test_records_and_bins_match_oracle(engine, name):
    targets, spec = SMALL[name]()
    orc = O.trace(targets, spec, use_bvh=False)
    recs, gbins, st = parity.run_gpu_records(engine, targets, spec)
    cmp = parity.compare_records(recs, orc, spec, name)
    parity.assert_records_equal(cmp)
    for k in ("segments", "hits", "shaded_hits", "refracted"):
        assert st[k] == orc["stats"][k], k
    n_out = parity.assert_bins_match(engine, targets, spec, orc, use_bvh=False)      # window-edge rays left out of both sides
    assert abs(int(st["captured"]) - int(orc["stats"]["captured"])) <= n_out
    assert engine.check_bvh() == 0


def test_empty_scene_and_direct_rays(engine):
    """No geometry at all: every ray is a direct ray; path rows stay -1; one direct bin per receiver."""
    import math
    from rts_b200.abi import PulseSpec
    spec = PulseSpec(grid=(1, 64, 64), max_refl=2, max_refr=0, tx_span=(0.2, 0.2, 0.0),
                     rx=[L.rx_sphere_from_desc((200.0, 0, 0), math.pi, 0.0, 10.0, 2.0, 2.0)], targ_vel=np.zeros((0, 3)))
    orc = O.trace([], spec)
    recs, gbins, st = parity.run_gpu_records(engine, [], spec)
    parity.assert_records_equal(parity.compare_records(recs, orc, spec, "empty"))
    assert len(gbins) == 1 and gbins[0]["direct"] == 1 and gbins[0]["npath"] == st["captured"] > 0


def test_single_ray_launch(engine):
    """d_width == 1: one ray along the boresight (ray_tracer.cu:160-163)."""
    targets, spec = scenes.flat_plate(n=1)
    orc = O.trace(targets, spec)
    recs, _, _ = parity.run_gpu_records(engine, targets, spec)
    parity.assert_records_equal(parity.compare_records(recs, orc, spec, "single"))
    assert recs[0]["reflDepth"][0] == 1


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz"))))
def test_gpu_matches_reference_golden_vectors(engine, path):
    """Golden vectors generated from the reference's own sources (tests/golden/make_golden.py)."""
    g = np.load(path)
    targets, spec = CASES[str(g["case"])](int(g["n"]))
    (res, ti, rcs, tp), _, _ = parity.run_gpu_records(engine, targets, spec)
    for f in ("reflDepth", "refrDepth", "rayLength", "firstHitPoint", "prevHitPoint", "power", "doppler", "received"):
        assert np.ascontiguousarray(res[f]).tobytes() == np.ascontiguousarray(g["results"][f]).tobytes(), f
    assert np.array_equal(ti, g["targ_intersect"]) and np.array_equal(tp, g["tri_path"])
    assert np.abs(rcs - g["rcs_angle"]).max() <= parity.RCS_ATOL


def test_leaf_bounds_are_the_reference_bound_program(engine):
    """triangle_mesh.cu:204-233: fp64 min/max narrowed with round-down / round-up."""
    targets, _ = scenes.ship(n=8, hull_res=24)
    engine.set_targets(targets)
    got = engine.tri_bounds()
    off = 0
    for t in targets:
        v = t.verts[t.tris]
        lo, hi = v.min(axis=1), v.max(axis=1)
        lo32, hi32 = lo.astype(np.float32), hi.astype(np.float32)
        lo32 = np.where(lo32.astype(np.float64) > lo, np.nextafter(lo32, np.float32(-np.inf)), lo32)
        hi32 = np.where(hi32.astype(np.float64) < hi, np.nextafter(hi32, np.float32(np.inf)), hi32)
        assert np.array_equal(got[off:off + len(t.tris), :3], lo32) and np.array_equal(got[off:off + len(t.tris), 3:], hi32)
        off += len(t.tris)
    assert engine.check_bvh() == 0


def test_moving_targets_refit_equals_rebuild_and_oracle(engine):
    """Per-pulse rigid motion on the device + BVH refit (replaces ray_tracer.cpp:936-1133): world vertices
    bit-equal to the host evaluation, hit ids identical between refit and fresh rebuild, and equal to the oracle."""
    ms = scenes.terrain_scene(n=160, cells_x=80, cells_y=40, movers=8, n_rx=2)
    engine.set_targets(ms.base)
    for pulse in (0, 1, 7, 40):
        rots, trans = ms.poses(pulse)
        engine.set_poses(rots, trans)
        world = ms.world_targets(pulse)
        for k in range(len(world)):
            assert engine.world_vertices(k).tobytes() == world[k].verts.tobytes(), (pulse, k)
        assert engine.check_bvh() == 0
        spec = ms.spec_for(pulse)
        engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
        refit = engine.records()
        refit_bins = engine.bins()
        engine.rebuild()
        engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
        rebuilt = engine.records()
        assert np.array_equal(refit[3], rebuilt[3]) and refit[0].tobytes() == rebuilt[0].tobytes()
        orc = O.trace(world, spec, use_bvh=True)
        parity.assert_records_equal(parity.compare_records(refit, orc, spec, f"pulse{pulse}"))
        obins, _ = O.trace_bins(world, spec, use_bvh=True)
        parity.assert_bins_close(parity.compare_bins(refit_bins, obins))
    # pulse 0 equals the static scene
    engine.set_targets(ms.world_targets(0))
    engine.trace(ms.spec_for(0), L.RTS_OUT_RECORDS)
    static0 = engine.records()
    engine.set_targets(ms.base)
    engine.set_poses(*ms.poses(0))
    engine.trace(ms.spec_for(0), L.RTS_OUT_RECORDS)
    moved0 = engine.records()
    assert static0[0].tobytes() == moved0[0].tobytes() and np.array_equal(static0[3], moved0[3])


def test_ray_shards_merge_to_the_full_launch(engine):
    """Ray sharding: bins of the shards, reduced like the NCCL exchange does, equal the single launch."""
    from rts_b200 import dist as rdist
    targets, spec = scenes.trihedral(n=256)
    engine.set_targets(targets)
    engine.trace(spec, L.RTS_OUT_BINS)
    full = engine.bins()
    world = 3
    parts = []
    import ctypes as C
    import torch
    for r in range(world):
        spec.ray_begin, spec.ray_count = rdist.shard_range(spec.rays, r, world)
        engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_FINALISE)
        sums, mins = rdist.bins_as_tensors(engine, torch.device("cuda:0"))
        parts.append((sums.cpu().numpy().reshape(-1, 5).copy(), mins.cpu().numpy().view(np.uint64).copy()))
    msum, mmin = rdist.merge_bins_numpy(parts)
    live = msum[:, 0] > 0
    assert live.sum() == len(full)
    assert np.array_equal(np.sort(msum[live, 0]), np.sort(full["npath"]))
    assert np.array_equal(np.sort(mmin[live]), np.sort(full["min_slot"]))
    assert np.isclose(msum[live, 1].sum(), full["sum_sqrt_power"].sum(), rtol=1e-12)
    spec.ray_begin = spec.ray_count = 0


def test_stride_sampling_matches_oracle(engine):
    targets, spec = scenes.trihedral(n=200)
    spec.ray_stride = 9
    engine.set_targets(targets)
    st = engine.trace(spec, L.RTS_OUT_BINS)
    obins, ost = O.trace_bins(targets, spec, use_bvh=False)
    assert st["primary_rays"] == ost["primary_rays"] and st["segments"] == ost["segments"]
    parity.assert_bins_close(parity.compare_bins(engine.bins(), obins))


# ---- BASELINE.json configs at (or near) full size ------------------------------------------------

def test_C2_trihedral_1M_rays(engine):
    targets, spec = scenes.trihedral(n=1000)
    engine.set_targets(targets)
    st = engine.trace(spec, L.RTS_OUT_BINS)
    obins, ost = O.trace_bins(targets, spec, use_bvh=False)
    for k in ("primary_rays", "segments", "hits", "shaded_hits", "captured"):
        assert st[k] == ost[k], k
    parity.assert_bins_close(parity.compare_bins(engine.bins(), obins))


def test_C3_ship_100k_triangles_refraction(engine):
    targets, spec = scenes.ship(n=768, hull_res=200)
    assert sum(len(t.tris) for t in targets) > 90000
    engine.set_targets(targets)
    assert engine.check_bvh() == 0
    st = engine.trace(spec, L.RTS_OUT_BINS | L.RTS_OUT_RECORDS)
    orc = O.trace(targets, spec, use_bvh=True)
    cmp = parity.compare_records(engine.records(), orc, spec, "C3")
    parity.assert_records_equal(cmp)
    for k in ("segments", "hits", "shaded_hits", "refracted"):
        assert st[k] == orc["stats"][k], k
    assert st["refracted"] > 10000
    parity.assert_bins_match(engine, targets, spec, orc, use_bvh=True)


def test_C4_terrain_1M_triangles(engine):
    ms = scenes.terrain_scene(n=1024, n_rx=8)
    assert sum(len(t.tris) for t in ms.base) >= 1_000_000
    engine.set_targets(ms.base)
    pulse = 3
    engine.set_poses(*ms.poses(pulse))
    assert engine.check_bvh() == 0
    spec = ms.spec_for(pulse)
    st = engine.trace(spec, L.RTS_OUT_BINS)
    world = ms.world_targets(pulse)
    orc = O.trace_shard(world, spec, use_bvh=True, arrays=False)      # counters + per-ray edge flags, no records
    ost = orc["stats"]
    assert st["primary_rays"] == orc["n_shard"] == 1024 * 1024
    for k in ("segments", "hits", "shaded_hits"):
        assert st[k] == ost[k], k
    # the oracle's window-edge rays (fp32 atan2f, libdevice vs glibc) are left out of both sides' bins; only they may flip a capture
    n_out = parity.assert_bins_match(engine, world, spec, orc, use_bvh=True, max_flagged=16)
    assert abs(int(st["captured"]) - int(ost["captured"])) <= n_out
    # size-independent properties at the full 16.7M-ray grid
    spec_full = scenes.terrain_scene(n=4096, cells_x=8, cells_y=4, movers=0).spec
    spec_full.rx, spec_full.targ_vel = spec.rx, spec.targ_vel
    st_full = engine.trace(spec_full, L.RTS_OUT_BINS)
    bins_full = engine.bins()
    assert st_full["primary_rays"] == 4096 * 4096 and st_full["segments"] >= st_full["primary_rays"]
    assert st_full["hits"] <= st_full["segments"] <= 4 * st_full["primary_rays"]
    assert int(bins_full["npath"][bins_full["direct"] == 0].sum() + (bins_full["npath"][bins_full["direct"] == 1].sum() if False else 0)) <= st_full["captured"]
    st_again = engine.trace(spec_full, L.RTS_OUT_BINS)
    for k in ("segments", "hits", "shaded_hits", "captured"):
        assert st_again[k] == st_full[k], k           # counts are deterministic
    again = engine.bins()
    assert np.array_equal(again["npath"], bins_full["npath"]) and np.array_equal(again["min_slot"], bins_full["min_slot"])
    assert np.allclose(again["sum_sqrt_power"], bins_full["sum_sqrt_power"], rtol=1e-12)


def test_async_pulses_equal_synchronous(engine):
    """RTS_ASYNC: pulses and pose updates are only enqueued; getters / rts_sync wait.  Same bins and stats as the
    synchronous calls, pulse after pulse (pinned staging ring reuse included: more pulses than ring slots)."""
    ms = scenes.terrain_scene(n=96, cells_x=80, cells_y=40, movers=6, n_rx=2)
    engine.set_targets(ms.base)
    want = []
    for pulse in range(12):
        engine.set_poses(*ms.poses(pulse))
        st = engine.trace(ms.spec_for(pulse), L.RTS_OUT_BINS)
        want.append((st, engine.bins().copy()))
    engine.set_targets(ms.base)
    for pulse in range(12):
        engine.set_poses(*ms.poses(pulse))
        assert engine.trace(ms.spec_for(pulse), L.RTS_OUT_BINS | L.RTS_ASYNC) is None
        if pulse % 3 == 2:          # read only some of them back; the others are overtaken by the next pulse
            got = engine.bins()
            st = engine.stats()
            for k in ("primary_rays", "segments", "hits", "shaded_hits", "captured"):
                assert st[k] == want[pulse][0][k], (pulse, k)
            parity.assert_bins_close(parity.compare_bins(got, want[pulse][1]), rtol=1e-12)
    engine.sync()


def test_targets_that_start_moving_later_and_drift_rebuild(engine):
    """Partial refit bookkeeping: a target joins the moving set at a later pulse, a pose repeated verbatim is a
    no-op, and a large displacement triggers the SAH-drift rebuild; hits always equal a fresh rebuild's."""
    ms = scenes.terrain_scene(n=128, cells_x=80, cells_y=40, movers=6, n_rx=1)
    engine.set_targets(ms.world_targets(0))             # committed at the pulse-0 poses: nothing moves at first
    K = len(ms.base)
    ident = ([None] * K, [(0.0, 0.0, 0.0)] * K)
    spec = ms.spec_for(0)

    from rts_b200.abi import Target
    committed = ms.world_targets(0)

    def hits_now(trans):
        """records after refit == records after a fresh rebuild == oracle on the translated meshes"""
        engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_NO_RCS_ANGLES)
        a = engine.records(rcs=False)
        engine.rebuild()
        engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_NO_RCS_ANGLES)
        b = engine.records(rcs=False)
        assert np.array_equal(a[3], b[3]) and a[0].tobytes() == b[0].tobytes()
        assert engine.check_bvh() == 0
        world = [Target(t.verts + np.asarray(d), t.tris, t.normals, t.refl_coeff, t.refr_index) for t, d in zip(committed, trans)]
        orc = O.trace(world, spec, use_bvh=True)
        assert np.array_equal(a[3], orc["tri_path"]) and np.array_equal(a[1], orc["targ_intersect"])
        return a

    engine.set_poses(*ident)                            # no-op
    hits_now(ident[1])
    trans = [list(t) for t in ident[1]]
    trans[2] = (15.0, -4.0, 6.0)                        # one mover starts moving
    engine.set_poses(ident[0], trans)
    hits_now(trans)
    trans[4] = (-9.0, 12.0, 3.0)                        # a second one joins later
    engine.set_poses(ident[0], trans)
    hits_now(trans)
    engine.set_poses(ident[0], trans)                   # verbatim repeat
    builds = engine.bvh_info().builds
    trans[2] = (2500.0, 900.0, 400.0)                   # far away: the refitted tree degrades, a rebuild follows
    engine.set_poses(ident[0], trans)
    engine.set_poses(ident[0], trans)
    trans[2] = (2501.0, 900.0, 400.0)
    engine.set_poses(ident[0], trans)
    hits_now(trans)
    assert engine.bvh_info().builds > builds


@pytest.mark.parametrize("name", ["plate", "trihedral", "slab", "terrain"])
def test_projected_primary_wave_is_bit_identical(engine, name, monkeypatch):
    """rts_b200/csrc/raster.cuh (default for nx == 1 launches; RTS_NO_RASTER=1 switches it off): primary visibility by
    projecting the triangles into the launch grid instead of walking the BVH per ray — same closest hits, hence the
    same records bit for bit, also for a shard."""
    if name == "plate":
        targets, spec = scenes.flat_plate(n=256)
    elif name == "trihedral":
        targets, spec = scenes.trihedral(n=300)
    elif name == "slab":
        targets, spec = scenes.slab(n=96)
    else:
        ms = scenes.terrain_scene(n=256, cells_x=100, cells_y=50, movers=6, n_rx=2)
        targets, spec = ms.world_targets(3), ms.spec_for(3)
    engine.set_targets(targets)
    for shard in ((0, 0, 0), (1, 0, 4)):
        spec.ray_begin, spec.ray_count, spec.ray_stride = shard
        engine.set_option("no_raster", 1)
        try:
            st0 = engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
            a, bins_a = engine.records(), engine.bins()
        finally:
            engine.set_option("no_raster", 0)
        st1 = engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
        b, bins_b = engine.records(), engine.bins()
        for k in ("segments", "hits", "shaded_hits", "captured", "edge_rays"):
            assert st0[k] == st1[k], (name, shard, k)
        assert st0["primary_projected"] == 0 and st1["primary_projected"] == 1
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[1], b[1])
        for f in a[0].dtype.names:
            assert a[0][f].tobytes() == b[0][f].tobytes(), (name, shard, f)
        parity.assert_bins_close(parity.compare_bins(bins_b, bins_a), rtol=1e-12)
    spec.ray_begin = spec.ray_count = spec.ray_stride = 0


def _raw_bins(engine):
    import torch
    from rts_b200 import dist as rdist
    sums, mins = rdist.bins_as_tensors(engine, torch.device("cuda:0"))
    return sums.cpu().numpy().reshape(-1, 5).copy(), mins.cpu().numpy().view(np.uint64).copy()


def test_C5_multistatic_eight_shards_and_batched_launch(engine):
    """C5 shape at reduced ray count: 1M-triangle terrain + movers, 8 receivers, rays dealt round-robin to 8 shards
    (what the 8 ranks of bench.py trace) — the reduced bins equal the single launch's.  Then a launch larger than
    one 2^24-ray batch: equal to the sum of its two halves traced separately."""
    from rts_b200 import dist as rdist
    ms = scenes.terrain_scene(n=2048, n_rx=8)
    engine.set_targets(ms.base)
    pulse = 5
    engine.set_poses(*ms.poses(pulse))
    spec = ms.spec_for(pulse)
    st_full = engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_FINALISE)
    full = _raw_bins(engine)
    parts, seg = [], 0
    for r in range(8):
        spec.ray_begin, spec.ray_count, spec.ray_stride = r, 0, 8
        st = engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_FINALISE)
        seg += st["segments"]
        parts.append(_raw_bins(engine))
    msum, mmin = rdist.merge_bins_numpy(parts)
    assert seg == st_full["segments"]
    assert np.array_equal(msum[:, 0], full[0][:, 0]) and np.array_equal(mmin, full[1])
    assert np.allclose(msum, full[0], rtol=1e-11, atol=0)
    assert (full[0][:, 0] > 0).sum() >= 8          # every receiver got something
    # batching: (1, 8192, 4096) = 2 x 2^24 primaries
    big = ms.spec_for(pulse)
    big.grid = (1, 8192, 4096)
    big.ray_begin = big.ray_count = big.ray_stride = 0
    st_big = engine.trace(big, L.RTS_OUT_BINS | L.RTS_NO_FINALISE)
    whole = _raw_bins(engine)
    halves, seg = [], 0
    for h in range(2):
        big.ray_begin, big.ray_count = h << 24, 1 << 24
        st = engine.trace(big, L.RTS_OUT_BINS | L.RTS_NO_FINALISE)
        seg += st["segments"]
        halves.append(_raw_bins(engine))
    hsum, hmin = rdist.merge_bins_numpy(halves)
    assert st_big["primary_rays"] == 2 << 24 and seg == st_big["segments"]
    assert np.array_equal(hsum[:, 0], whole[0][:, 0]) and np.array_equal(hmin, whole[1])
    assert np.allclose(hsum, whole[0], rtol=1e-11, atol=0)


def test_builder_choice_and_ploc_parity(monkeypatch):
    """Two topology builders (Morton radix tree, PLOC) feed the same refit / traversal code; a scene keeps the one with
    the lower SAH cost.  Forced PLOC reproduces the oracle bit for bit, including moving targets with partial refit."""
    with L.Engine(0) as eng:                                     # automatic choice
        targets, spec = scenes.ship(n=64, hull_res=64)
        eng.set_targets(targets)
        assert eng.bvh_info().builder == 2                       # irregular mesh + huge sea triangles: PLOC
        ms = scenes.terrain_scene(n=64, cells_x=200, cells_y=100, movers=0)
        eng.set_targets(ms.base)
        assert eng.bvh_info().builder == 1                       # regular height field: the radix tree
    monkeypatch.setenv("RTS_BVH", "ploc")
    with L.Engine(0) as eng:
        for name in ("slab", "ship_small", "spheres_interp"):
            targets, spec = SMALL[name]()
            orc = O.trace(targets, spec, use_bvh=False)
            recs, gbins, st = parity.run_gpu_records(eng, targets, spec)
            assert eng.bvh_info().builder == 2 and eng.check_bvh() == 0
            parity.assert_records_equal(parity.compare_records(recs, orc, spec, name + "/ploc"))
        ms = scenes.terrain_scene(n=128, cells_x=80, cells_y=40, movers=8, n_rx=2)
        eng.set_targets(ms.base)
        for pulse in (0, 3, 9):
            eng.set_poses(*ms.poses(pulse))
            spec = ms.spec_for(pulse)
            eng.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
            recs = eng.records()
            assert eng.check_bvh() == 0 and eng.bvh_info().builder == 2
            orc = O.trace(ms.world_targets(pulse), spec, use_bvh=True)
            parity.assert_records_equal(parity.compare_records(recs, orc, spec, f"ploc/pulse{pulse}"))


def _shifted(targets, spec, shift):
    from rts_b200.abi import Target
    shift = np.asarray(shift, dtype=np.float64)
    t2 = [Target(t.verts + shift, t.tris, t.normals, t.refl_coeff, t.refr_index) for t in targets]
    spec.tx_origin = tuple(np.asarray(spec.tx_origin) + shift)
    for r in spec.rx:
        for a in range(3):
            r.centre[a] += float(shift[a])
    return t2, spec


@pytest.mark.parametrize("name", ["earth_scale", "degenerate", "axis_parallel"])
def test_conservative_traversal_edge_cases(engine, name):
    """The fp32 slab test must never prune a box the fp64 triangle test could hit: coordinates at Earth-radius scale
    (the reference hard-codes an Earth sphere, ray_tracer.cu:447, so such scenes are legal; fp32 ulp there is 0.5 m),
    zero-area triangles (n.d = 0 -> NaN, rejected by the comparisons, triangle_mesh.cu:124-136), and rays with
    exactly-zero direction components (odd grid: the centre ray runs along the boresight)."""
    if name == "earth_scale":
        targets, spec = scenes.trihedral(n=120)
        targets, spec = _shifted(targets, spec, (6.3e6, 1.0e5, -2.0e5))
    elif name == "degenerate":
        from rts_b200.abi import Target
        targets, spec = scenes.slab(n=64)
        t = targets[0]
        v = np.vstack([t.verts, [[49.0, 0.0, 0.0], [49.0, 1.0, 0.0], [49.0, 2.0, 0.0], [49.5, 0.3, 0.2]]])
        nv = len(t.verts)
        tris = np.vstack([t.tris, [[nv, nv + 1, nv + 2], [nv + 3, nv + 3, nv + 3]]]).astype(np.uint32)   # collinear, and a point
        nrm = np.vstack([t.normals, [[1.0, 0, 0], [1.0, 0, 0]]])
        targets = [Target(v, tris, nrm, t.refl_coeff, t.refr_index)]
    else:
        targets, spec = scenes.flat_plate(n=65)
        spec.max_refl = 2
    orc = O.trace(targets, spec, use_bvh=False)
    recs, gbins, st = parity.run_gpu_records(engine, targets, spec)
    cmp = parity.compare_records(recs, orc, spec, name)
    parity.assert_records_equal(cmp)
    for k in ("segments", "hits", "shaded_hits"):
        assert st[k] == orc["stats"][k], k
    assert st["hits"] > 0 and engine.check_bvh() == 0


def test_primary_direction_buffer_reuse(engine):
    """The projected primary wave keeps the ray directions of the previous pulse when the launch geometry is the same
    and regenerates them when the boresight, span, grid or shard changes: records equal the oracle every time."""
    targets, spec = scenes.trihedral(n=160)
    engine.set_targets(targets)
    import copy
    variants = []
    for k, (daz, n, stride) in enumerate([(0.0, 160, 0), (0.0, 160, 0), (0.01, 160, 0), (0.0, 160, 0), (0.0, 128, 0), (0.0, 128, 3), (0.0, 128, 3)]):
        s = copy.copy(spec)
        s.tx_dir = (spec.tx_dir[0] + daz, spec.tx_dir[1])
        s.grid = (1, n, n)
        s.ray_begin, s.ray_count, s.ray_stride = (1 if stride else 0), 0, stride
        variants.append(s)
    for k, s in enumerate(variants):
        st = engine.trace(s, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
        assert st["primary_projected"] == 1
        if s.ray_stride:      # the oracle leaves the slots of other shards untouched: compare what the shard produced
            obins, ost = O.trace_bins(targets, s, use_bvh=False)
            assert st["segments"] == ost["segments"] and st["hits"] == ost["hits"]
            parity.assert_bins_close(parity.compare_bins(engine.bins(), obins))
        else:
            orc = O.trace(targets, s, use_bvh=False)
            parity.assert_records_equal(parity.compare_records(engine.records(), orc, s, f"variant{k}"))


def test_kept_hits_over_consecutive_pulses_match_the_oracle(engine):
    """coherent.cuh / raster.cuh keep the static part of the primary hits and of the first reflections across pulses with
    the same launch.  Consecutive pulses (no rebuild in between) against the oracle: movers low over the terrain so that
    first reflections do run into them, a shading switch changing mid-way (refill), and a receiver change (no refill)."""
    import copy
    ms = scenes.terrain_scene(n=192, cells_x=80, cells_y=40, movers=8, n_rx=2)
    ms.positions0[1:, 2] = 18.0 + 4.0 * np.arange(len(ms.positions0) - 1)     # just above the hills
    ms.velocities[1:] *= 40.0                                                   # metres per pulse, so pixels change hands
    engine.set_targets(ms.base)
    seen_kept = 0
    for k, pulse in enumerate([0, 1, 2, 3, 4, 5, 6, 7]):
        engine.set_poses(*ms.poses(pulse))
        spec = ms.spec_for(pulse)
        if k >= 4:
            spec.max_refl = 2                       # different depth budget: the kept reflections are refilled
        if k >= 6:
            spec.rx = list(reversed(spec.rx))       # receivers do not enter the kept data
        st = engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
        assert st["primary_projected"] == 1
        world = ms.world_targets(pulse)
        orc = O.trace(world, spec, use_bvh=True)
        cmp = parity.compare_records(engine.records(), orc, spec, f"pulse{pulse}")
        parity.assert_records_equal(cmp)
        for key in ("segments", "hits", "shaded_hits"):
            assert st[key] == orc["stats"][key], (pulse, key)
        flagged = np.nonzero((np.asarray(orc["edge"]) & O.EDGE_WINDOW) != 0)[0]
        if len(flagged) == 0:       # the bins of the very pulse that used the kept hits
            obins, _ = O.trace_bins(world, spec, use_bvh=True)
            parity.assert_bins_close(parity.compare_bins(engine.bins(), obins))
        else:                       # window-edge rays left out of both sides — on a second engine: sub-shard launches would reset what this one keeps
            aux = L.Engine(0)
            try:
                aux.set_targets(world)
                parity.assert_bins_match(aux, world, spec, orc, use_bvh=True)
            finally:
                aux.close()
        seen_kept += 1 if st["kept_reflections"] > 0 else 0
        if k in (1, 2, 3, 5, 6, 7):
            assert 0 < st["kept_reflections"] < st["segments"] - st["primary_rays"]   # most, but not all: movers in the way
    assert seen_kept >= 7 and engine.bvh_info().builds >= 2


def test_no_reuse_flag_traces_from_scratch(engine):
    """RTS_NO_REUSE: nothing kept from earlier pulses is used (bench.py's headline legs); identical bins either way."""
    ms = scenes.terrain_scene(n=160, cells_x=80, cells_y=40, movers=6, n_rx=2)
    engine.set_targets(ms.base)
    for pulse in range(4):
        engine.set_poses(*ms.poses(pulse))
        spec = ms.spec_for(pulse)
        a = engine.trace(spec, L.RTS_OUT_BINS)
        bins_a = engine.bins().copy()
        b = engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_REUSE)
        bins_b = engine.bins().copy()
        assert b["kept_reflections"] == 0 and (pulse == 0 or a["kept_reflections"] > 0)
        for k in ("segments", "hits", "shaded_hits", "captured"):
            assert a[k] == b[k], (pulse, k)
        parity.assert_bins_close(parity.compare_bins(bins_a, bins_b), rtol=1e-12)
