"""Randomised parity: triangle soups with mixed materials, normal conventions, launch shapes, receivers and target
velocities, traced through the C-ABI and bit-compared (records) with the exhaustive-search oracle — the definitional
closest hit, independent of any BVH.  Seeds are fixed; every case is a few thousand rays."""
import math

import numpy as np
import pytest

import oracle_api as O
import parity
from rts_b200 import lib as L
from rts_b200 import scenes
from rts_b200.abi import PulseSpec, Target

pytestmark = pytest.mark.gpu


from soups import case as _case  # noqa: E402


@pytest.mark.parametrize("seed", range(24))
def test_random_soups_match_the_exhaustive_oracle(engine, seed):
    targets, spec = _case(seed)
    orc = O.trace(targets, spec, use_bvh=False)
    recs, gbins, st = parity.run_gpu_records(engine, targets, spec)
    cmp = parity.compare_records(recs, orc, spec, f"fuzz/{seed}")
    parity.assert_records_equal(cmp)
    for k in ("segments", "hits", "shaded_hits", "refracted"):
        assert st[k] == orc["stats"][k], (k, st[k], orc["stats"][k])
    obins, _ = O.trace_bins(targets, spec, use_bvh=False)
    parity.assert_bins_close(parity.compare_bins(gbins, obins))
    assert engine.check_bvh() == 0
    # the same launch through the BVH primary wave instead of the projection must give the same records
    if spec.grid[0] == 1:
        engine.set_option("no_raster", 1)
        try:
            recs2, gbins2, st2 = parity.run_gpu_records(engine, targets, spec)
        finally:
            engine.set_option("no_raster", 0)
        parity.assert_records_equal(parity.compare_records(recs2, orc, spec, f"fuzz/{seed}/bvh"))
        assert st2["primary_projected"] == 0 and st["primary_projected"] == 1
    # the launch cut into batches of an odd size (RTS_BATCH: the 2^24-ray batching of large launches at test size)
    engine.set_option("batch", 257 + 64 * seed)
    try:
        recs3, gbins3, st3 = parity.run_gpu_records(engine, targets, spec)
    finally:
        engine.set_option("batch", 0)
    parity.assert_records_equal(parity.compare_records(recs3, orc, spec, f"fuzz/{seed}/batched"))
    parity.assert_bins_close(parity.compare_bins(gbins3, obins))
    assert st3["segments"] == st["segments"] and st3["hits"] == st["hits"]
    # every later wave through the two-kernel form (split.cuh: k_traverse + k_shade_wave), which large waves take by
    # themselves: records mode keeps every ray, bins-only mode drops the rays that can no longer change an output
    engine.set_option("split_below", 0)
    try:
        recs4, gbins4, st4 = parity.run_gpu_records(engine, targets, spec)
        st5 = engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_REUSE)
        gbins5 = engine.bins()
        engine.set_option("no_raster", 1)
        st6 = engine.trace(spec, L.RTS_OUT_BINS | L.RTS_COUNT_NODES)
        gbins6 = engine.bins()
    finally:
        engine.set_option("split_below", 1 << 18)
        engine.set_option("no_raster", 0)
    parity.assert_records_equal(parity.compare_records(recs4, orc, spec, f"fuzz/{seed}/split"))
    parity.assert_bins_close(parity.compare_bins(gbins4, obins))
    parity.assert_bins_close(parity.compare_bins(gbins5, obins))
    parity.assert_bins_close(parity.compare_bins(gbins6, obins))
    for k in ("segments", "hits", "shaded_hits", "refracted", "captured"):
        assert st4[k] == st[k] and st5[k] == st[k] and st6[k] == st[k], (k, st[k], st4[k], st5[k], st6[k])


@pytest.mark.parametrize("seed", range(6))
def test_random_moving_scenes_with_between_pulse_reuse(engine, seed):
    """Terrain + movers from different PRNG seeds, pulses visited out of order and repeated, with the library's default
    between-pulse reuse (kept static hits and first reflections, partial refit): every pulse equals the oracle run on
    that pulse's world-space meshes."""
    rng = np.random.default_rng(77 + seed)
    ms = scenes.terrain_scene(n=int(rng.integers(72, 120)), cells_x=int(rng.integers(24, 64)), cells_y=int(rng.integers(12, 32)),
                              movers=int(rng.integers(1, 9)), n_rx=int(rng.integers(1, 4)), seed=0x52545301 + 17 * seed)
    engine.set_targets(ms.base)
    order = [0, 1, 1, 4, 2, 2, 7, 3]
    engine.set_option("split_below", 0 if seed % 2 else 1 << 18)    # odd seeds: later waves through k_traverse + k_shade_wave
    for i, p in enumerate(order):
        engine.set_poses(*ms.poses(p))
        spec = ms.spec_for(p)
        flags = L.RTS_OUT_BINS | (L.RTS_ASYNC if i % 2 else 0)
        engine.trace(spec, flags)
        st = engine.stats()
        obins, ost = O.trace_bins(ms.world_targets(p), spec, use_bvh=True)
        for k in ("segments", "hits", "shaded_hits", "captured"):
            assert st[k] == ost[k], (seed, i, p, k, st[k], ost[k])
        parity.assert_bins_close(parity.compare_bins(engine.bins(), obins))
    engine.set_option("split_below", 1 << 18)
    assert engine.check_bvh() == 0


@pytest.mark.parametrize("seed", range(12))
def test_random_shards_builders_and_changing_launches(seed, monkeypatch):
    """Random ray shards (begin / count / stride), forced topology builders and leaf sizes, and a transmitter that moves
    between pulses (so everything kept between pulses must be dropped): records and bins equal the oracle's for the
    same shard."""
    rng = np.random.default_rng(5000 + seed)
    monkeypatch.setenv("RTS_BVH", ["lbvh", "ploc"][seed % 2])
    monkeypatch.setenv("RTS_LEAF_MAX", str(int(rng.integers(1, 5))))
    targets, spec = _case(100 + seed)
    with L.Engine(0) as eng:
        eng.set_targets(targets)
        for pulse in range(3):
            s = PulseSpec(**{**spec.__dict__})
            s.ray_begin = int(rng.integers(0, spec.rays // 2))
            s.ray_count = int(rng.integers(0, spec.rays // 2)) if pulse != 1 else 0
            s.ray_stride = int(rng.choice([0, 1, 2, 3, 5, 8]))
            if pulse:
                s.tx_origin = tuple(np.asarray(spec.tx_origin) + rng.normal(0, 3.0, 3))
                s.tx_dir = (spec.tx_dir[0] + float(rng.normal(0, 0.05)), spec.tx_dir[1] + float(rng.normal(0, 0.05)))
            orc = O.trace(targets, s, use_bvh=False)
            st = eng.trace(s, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
            recs = eng.records()
            # rays outside the shard keep each side's buffer defaults: compare the shard's own result slots
            stride = max(1, s.ray_stride)
            end = min(s.rays, s.ray_begin + s.ray_count) if s.ray_count else s.rays
            sel = np.arange(s.ray_begin, end, stride)
            rows = np.concatenate([sel + k * s.rays for k in range(s.slots)])
            sub = PulseSpec(**{**s.__dict__})
            sub.grid = (1, 1, len(sel))
            orc_sel = {"edge": orc["edge"][sel], "results": orc["results"][rows], "tri_path": orc["tri_path"][rows],
                       "targ_intersect": orc["targ_intersect"][rows], "rcs_angle": orc["rcs_angle"][rows]}
            recs_sel = (recs[0][rows], recs[1][rows], recs[2][rows] if recs[2] is not None else None, recs[3][rows])
            assert len(sel) == st["primary_rays"]
            parity.assert_records_equal(parity.compare_records(recs_sel, orc_sel, sub, f"shards/{seed}/{pulse}"))
            for k in ("primary_rays", "segments", "hits", "shaded_hits", "refracted"):
                assert st[k] == orc["stats"][k], (k, st[k], orc["stats"][k])
            obins, _ = O.trace_bins(targets, s, use_bvh=False)
            parity.assert_bins_close(parity.compare_bins(eng.bins(), obins))
            assert eng.check_bvh() == 0


@pytest.mark.parametrize("seed", range(8))
def test_random_rigid_motion_of_soups(seed):
    """Soups centred on their own origins, moved per pulse by random translations and yaw/pitch/roll rates (device
    transform + partial refit, some targets only start moving at a later pulse, large motions force SAH-drift
    rebuilds): per-ray records of every pulse equal the exhaustive oracle on the host-transformed meshes."""
    rng = np.random.default_rng(9000 + seed)
    targets, spec = _case(200 + seed)
    K = len(targets)
    base, pos0 = [], np.zeros((K, 3))
    for k, t in enumerate(targets):
        c = t.verts.mean(axis=0)
        base.append(Target(t.verts - c, t.tris, t.normals, t.refl_coeff, t.refr_index))
        pos0[k] = c
    starts = rng.integers(0, 3, K)                       # pulse index from which the target moves
    vel = rng.normal(0.0, 2500.0, (K, 3))                # m/s; pri = 1 ms -> metres per pulse
    rates = rng.normal(0.0, 300.0, (K, 3)) * (rng.random((K, 1)) < 0.5)
    ms = scenes.MovingScene(base=base, spec=spec, positions0=pos0, velocities=vel, rot_rates=rates)
    with L.Engine(0) as eng:
        eng.set_targets(base)                            # committed at the origin; every pulse's pose translates to pos0 + ...
        for pulse in [0, 1, 2, 3, 5, 4]:
            rots, trans = ms.poses(pulse)
            world = ms.world_targets(pulse)
            for k in range(K):                            # targets that have not started yet stay at their pulse-0 pose
                if pulse < starts[k]:
                    rots[k], trans[k] = None, pos0[k]
                    world[k] = Target(base[k].verts + pos0[k][None, :], base[k].tris, base[k].normals, base[k].refl_coeff, base[k].refr_index)
            eng.set_poses(rots, trans)
            s = ms.spec_for(pulse)
            orc = O.trace(world, s, use_bvh=False)
            st = eng.trace(s, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
            parity.assert_records_equal(parity.compare_records(eng.records(), orc, s, f"motion/{seed}/{pulse}"))
            for k in ("segments", "hits", "shaded_hits", "refracted"):
                assert st[k] == orc["stats"][k], (k, pulse, st[k], orc["stats"][k])
            obins, _ = O.trace_bins(world, s, use_bvh=False)
            parity.assert_bins_close(parity.compare_bins(eng.bins(), obins))
            assert eng.check_bvh() == 0


@pytest.mark.parametrize("seed", range(10))
def test_random_post_process_and_responses(engine, seed):
    """Per-target RCS and antenna gains on random soups: the fused bins / rts_get_responses, the two-phase path
    (rts_get_received -> host callbacks -> rts_aggregate) and the oracle's literal host flow (post-process, O(R^2)
    aggregation, sort + unique: ray_tracer.cpp:1190-1320) all give the same responses."""
    rng = np.random.default_rng(4000 + seed)
    targets, spec = _case(300 + seed)
    for t in targets:                                   # powers stay positive so that sqrt(P) and the comparisons are meaningful
        t.refl_coeff = abs(t.refl_coeff) if t.refl_coeff != 0 else 0.5
    spec.targ_rcs = rng.uniform(0.2, 9.0, len(targets))
    spec.gain_tx, spec.gain_rx = float(rng.uniform(1, 40)), float(rng.uniform(1, 40))
    engine.set_targets(targets)
    engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
    gbins, got = engine.bins(), engine.responses()
    obins, _ = O.trace_bins(targets, spec, use_bvh=False)
    parity.assert_bins_close(parity.compare_bins(gbins, obins))
    r = O.trace(targets, spec, use_bvh=False)
    rx_res, rx_rows, rx_slots = O.postprocess(r["results"], r["targ_intersect"], spec, rcs_per_target=spec.targ_rcs,
                                              gain=spec.gain_tx * spec.gain_rx)
    if len(rx_res) == 0:
        assert len(got) == 0 and len(gbins) == 0
        return
    want = O.responses(O.aggregate(rx_res, rx_rows, spec, literal=True), rx_slots)
    assert len(got) == len(want)
    assert np.array_equal(got["rx"], want["rx"]) and np.array_equal(got["slot"], want["slot"])
    for f in ("power", "delay", "doppler", "phase"):
        assert np.allclose(got[f], want[f], rtol=1e-5, atol=1e-300), f
    # two-phase: the received rays compacted on the device, the callbacks on the host, the aggregation on the device
    d_res, d_ti, _, d_slots = engine.received()
    assert np.array_equal(d_slots, rx_slots.astype(np.uint64))
    a = engine.aggregate(rx_res, rx_rows, spec.cspeed, spec.carrier, ray_total=spec.ray_total)
    two = O.responses(a, rx_slots)
    assert np.array_equal(two["slot"], want["slot"])
    for f in ("power", "delay", "doppler", "phase"):
        assert np.allclose(two[f], want[f], rtol=1e-5, atol=1e-300), f
