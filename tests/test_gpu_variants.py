"""The alternative forms of the bounce waves kept behind options (rts_set_option) give the same answers as the default:
split later waves (k_traverse + k_shade_wave over the quantised nodes), primary shading without the in-place first
reflection (no_follow), global-atomics-only bins (no_smem_bins), BVH primary wave (no_raster), everything on one stream
(no_overlap: no side streams for the direction pass and the refit), never-moving and moving triangles projected in one
pass (no_split_raster 1) or always in two (-1; by default in two only while an earlier pulse or batch is in flight).  Records bit-equal, bins
equal in the exact fields and to 1e-12 in the sums (fp64 atomics commute, their rounding order does not)."""
import numpy as np
import pytest

from rts_b200 import lib as L, scenes

pytestmark = pytest.mark.gpu

EXACT = ("rx", "path", "npath", "min_slot", "own_min_slot", "direct")
SUMS = ("sum_sqrt_power", "sum_delay", "sum_phase", "sum_doppler", "power", "delay", "phase", "doppler")


def _run(engine, spec, flags):
    st = engine.trace(spec, flags | L.RTS_NO_REUSE)
    return st, engine.bins().copy()


@pytest.mark.parametrize("option,value", [("no_split", 0), ("no_follow", 1), ("no_smem_bins", 1), ("no_raster", 1), ("no_overlap", 1), ("no_split_raster", 1), ("no_split_raster", -1)])
@pytest.mark.parametrize("scene", ["terrain", "trihedral", "slab"])
def test_option_gives_the_default_answers(engine, option, value, scene):
    if scene == "terrain":
        ms = scenes.terrain_scene(n=512, cells_x=96, cells_y=48, movers=6, n_rx=2)
        targets, spec, poses = ms.base, ms.spec_for(3), ms.poses(3)
    elif scene == "trihedral":
        (targets, spec), poses = scenes.trihedral(n=400), None
    else:
        (targets, spec), poses = scenes.slab(n=160), None          # refraction: both ends of the queue in use
    engine.set_targets(targets)
    if poses is not None:
        engine.set_poses(*poses)
    st0, bins0 = _run(engine, spec, L.RTS_OUT_BINS)
    engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_NO_REUSE)
    rec0 = engine.records()
    default = {"no_split": 1, "no_follow": 0, "no_smem_bins": 0, "no_raster": 0, "no_overlap": 0, "no_split_raster": 0}[option]
    try:
        engine.set_option(option, value)
        if option == "no_split":
            engine.set_option("split_below", 1024)                  # the test scenes are small
            if poses is not None:
                engine.set_poses(*poses)                            # enabling the split form rebuilt the tree at the current poses
        st1, bins1 = _run(engine, spec, L.RTS_OUT_BINS)
        engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_NO_REUSE)
        rec1 = engine.records()
    finally:
        engine.set_option(option, default)
        if option == "no_split":
            engine.set_option("split_below", 1 << 18)
    for k in ("segments", "hits", "shaded_hits", "captured", "refracted"):
        assert st0[k] == st1[k], k
    assert len(bins0) == len(bins1) > 0
    for f in EXACT:
        assert np.array_equal(bins0[f], bins1[f]), f
    for f in SUMS:
        assert np.allclose(bins0[f], bins1[f], rtol=1e-12, atol=0), f
    for a, b in zip(rec0, rec1):
        if a.dtype.names:
            for f in a.dtype.names:
                assert a[f].tobytes() == b[f].tobytes(), f
        else:
            assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("reuse", [False, True])
def test_pulses_enqueued_back_to_back_equal_pulses_traced_alone(engine, reuse):
    """Pulses of a moving scene enqueued without waiting (RTS_ASYNC): the next pulse's direction pass and static footprints
    run on the library's side stream beside the previous pulse's last waves, its refit beside the footprint kernels.  The
    bins of the last pulse of every prefix equal the bins of that pulse traced with nothing in flight, on one stream."""
    ms = scenes.terrain_scene(n=384, cells_x=96, cells_y=48, movers=6, n_rx=2)
    engine.set_targets(ms.base)
    flags = L.RTS_OUT_BINS | (0 if reuse else L.RTS_NO_REUSE)
    pulses = [0, 3, 1, 7, 2, 5]
    alone = {}
    engine.set_option("no_overlap", 1)
    try:
        for p in pulses:
            engine.set_poses(*ms.poses(p))
            st = engine.trace(ms.spec_for(p), flags | L.RTS_NO_REUSE)
            alone[p] = (st, engine.bins().copy())
    finally:
        engine.set_option("no_overlap", 0)
    for last in range(1, len(pulses) + 1):
        for p in pulses[:last]:
            engine.set_poses(*ms.poses(p))
            engine.trace(ms.spec_for(p), flags | L.RTS_ASYNC)
        st = engine.stats()                     # waits for the last pulse
        bins = engine.bins().copy()
        st0, bins0 = alone[pulses[last - 1]]
        for k in ("segments", "hits", "shaded_hits", "captured"):
            assert st[k] == st0[k], (last, k)
        assert len(bins) == len(bins0) > 0
        for f in EXACT:
            assert np.array_equal(bins[f], bins0[f]), (last, f)
        for f in SUMS:
            assert np.allclose(bins[f], bins0[f], rtol=1e-12, atol=0), (last, f)


def test_bins_of_the_previous_pulse_while_the_next_one_runs(engine):
    """rts_get_bins_previous: with pulse p+1 enqueued (RTS_ASYNC), pulse p's bins come from the pinned block pulse p filled;
    equal to the bins of pulse p traced alone.  Without a previous pulse's bins the call fails, it does not guess."""
    ms = scenes.terrain_scene(n=384, cells_x=96, cells_y=48, movers=6, n_rx=2)
    engine.set_targets(ms.base)
    flags = L.RTS_OUT_BINS | L.RTS_NO_REUSE
    pulses = [0, 3, 1, 7]
    alone = {}
    for p in pulses:
        engine.set_poses(*ms.poses(p))
        engine.trace(ms.spec_for(p), flags)
        alone[p] = engine.bins().copy()
    engine.trace(ms.spec_for(pulses[-1]), L.RTS_OUT_RECORDS)          # a pulse that leaves no bins
    engine.set_poses(*ms.poses(pulses[0]))
    engine.trace(ms.spec_for(pulses[0]), flags | L.RTS_ASYNC)
    with pytest.raises(RuntimeError):
        engine.bins_previous()
    got = []
    for i, p in enumerate(pulses[1:], 1):
        engine.set_poses(*ms.poses(p))
        engine.trace(ms.spec_for(p), flags | L.RTS_ASYNC)
        got.append((pulses[i - 1], engine.bins_previous().copy()))
    got.append((pulses[-1], engine.bins().copy()))
    for p, bins in got:
        assert len(bins) == len(alone[p]) > 0
        for f in EXACT:
            assert np.array_equal(bins[f], alone[p][f]), (p, f)
        for f in SUMS:
            assert np.allclose(bins[f], alone[p][f], rtol=1e-12, atol=0), (p, f)
