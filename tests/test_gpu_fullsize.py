"""BASELINE.json's configurations at their real launch grids, sampled with ray_stride so that the oracle finishes in
seconds: the rays of the sample are the very rays of the full launch (same launch indices, same directions, same scene),
traced through the C-ABI and compared with the oracle ray by ray (records of the shard, gathered on the device) and bin
by bin.  Rays the oracle flags as receiver-window edge cases (fp32 atan2f differs between libdevice and glibc by a few
ulp) are left out of BOTH sides' bins by cutting the shard around them — no comparison is skipped; the number of rays
left out is asserted to be small.  Reference semantics: ray_tracer.cu:144-255 (ray generation), :326-375 (capture)."""
import numpy as np
import pytest

import oracle_api as O
import parity
from rts_b200 import lib as L
from rts_b200 import scenes

pytestmark = pytest.mark.gpu


def _window_edge(orc):
    return np.nonzero((orc["edge"] & O.EDGE_WINDOW) != 0)[0]


def test_C3_ship_full_grid_stride_31(engine):
    """C3: 100k-triangle dielectric ship, (1,4096,4096) rays, refraction, 4 Rx — every 31st ray of the real grid."""
    targets, spec = scenes.ship(n=4096, hull_res=200)
    assert spec.grid == (1, 4096, 4096) and sum(len(t.tris) for t in targets) > 90000 and len(spec.rx) == 4
    spec.ray_begin, spec.ray_count, spec.ray_stride = 7, 0, 31
    engine.set_targets(targets)
    st = engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
    orc = O.trace_shard(targets, spec, use_bvh=True)
    assert st["primary_rays"] == orc["n_shard"] == O.shard_size(spec) > 500_000
    cmp = parity.compare_shard_records(engine.records_shard(), orc, spec, "C3/full-grid")
    parity.assert_records_equal(cmp)
    for k in ("segments", "hits", "shaded_hits", "refracted"):
        assert st[k] == orc["stats"][k], k
    assert st["refracted"] > 10000
    flagged = _window_edge(orc)
    assert len(flagged) <= 16, len(flagged)
    gb, ob, n_out = parity.bins_excluding(engine, targets, spec, flagged)
    parity.assert_bins_close(parity.compare_bins(gb, ob))
    assert len(gb) > 0 and n_out == len(flagged)


@pytest.mark.parametrize("pulse", [0, 3, 63])
def test_C4_bench_scene_full_grid_stride_16(engine, pulse):
    """C4 exactly as bench.py runs it (1,010,336 triangles, 16 movers, (1,4096,4096) rays, per-pulse poses + refit),
    every 16th ray: counters and bins against the oracle on that pulse's world-space meshes."""
    ms = scenes.terrain_scene(n=4096, n_rx=1, nz=4096)
    engine.set_targets(ms.base)
    engine.set_poses(*ms.poses(pulse))
    assert engine.check_bvh() == 0
    spec = ms.spec_for(pulse)
    spec.ray_begin, spec.ray_count, spec.ray_stride = pulse % 16, 0, 16
    world = ms.world_targets(pulse)
    orc = O.trace_shard(world, spec, use_bvh=True, arrays=False)
    st = engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_REUSE)
    assert st["primary_rays"] == orc["n_shard"] == 4096 * 4096 // 16
    for k in ("segments", "hits", "shaded_hits"):
        assert st[k] == orc["stats"][k], (pulse, k, st[k], orc["stats"][k])
    flagged = _window_edge(orc)
    assert len(flagged) <= 16, len(flagged)
    assert abs(int(st["captured"]) - int(orc["stats"]["captured"])) <= len(flagged)
    gb, ob, _ = parity.bins_excluding(engine, world, spec, flagged, gpu_flags=L.RTS_NO_REUSE)
    parity.assert_bins_close(parity.compare_bins(gb, ob))
    assert len(gb) > 0
    # the same pulse with the library's between-pulse reuse on (its default) gives the same bins
    gb2, _, _ = parity.bins_excluding(engine, world, spec, flagged)
    parity.assert_bins_close(parity.compare_bins(gb2, ob))


@pytest.mark.parametrize("rank", range(8))
def test_C5_multistatic_shard_of_a_1e8_ray_pulse(engine, rank):
    """C5: 8 receivers, (1,10000,10000) rays per pulse dealt round-robin to 8 ranks; every 64th ray of rank `rank`'s share."""
    ms = scenes.terrain_scene(n=10000, n_rx=8, nz=10000)
    pulse = 5
    engine.set_targets(ms.base)
    engine.set_poses(*ms.poses(pulse))
    spec = ms.spec_for(pulse)
    assert spec.rays == 10 ** 8 and len(spec.rx) == 8
    spec.ray_begin, spec.ray_count, spec.ray_stride = rank, 0, 8 * 64
    world = ms.world_targets(pulse)
    orc = O.trace_shard(world, spec, use_bvh=True, arrays=False)
    st = engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_REUSE)
    assert st["primary_rays"] == orc["n_shard"] == O.shard_size(spec)
    for k in ("segments", "hits", "shaded_hits"):
        assert st[k] == orc["stats"][k], (rank, k)
    flagged = _window_edge(orc)
    assert len(flagged) <= 16, len(flagged)
    gb, ob, _ = parity.bins_excluding(engine, world, spec, flagged, gpu_flags=L.RTS_NO_REUSE)
    parity.assert_bins_close(parity.compare_bins(gb, ob))
