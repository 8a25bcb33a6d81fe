"""Aggregation known answers (aggregation.cu:32-97, ray_tracer.cpp:1266-1321): the literal O(R^2)
transcription against the binned form and against hand-derived values."""
import numpy as np
import pytest

import oracle_api as O
from rts_b200.abi import RAY_RECORD, PulseSpec

SPEC = PulseSpec(grid=(1, 1, 1), max_refl=3, max_refr=0)


def make(n, rx, rows, length, power, doppler=0.0, refl=1):
    r = np.zeros(n, dtype=RAY_RECORD)
    r["received"] = rx
    r["rayLength"] = length
    r["power"] = power
    r["doppler"] = doppler
    r["reflDepth"] = refl
    return r, np.asarray(rows, dtype=np.int32).reshape(n, -1)


def test_identical_rays_one_path():
    res, rows = make(7, 0, [[2, -1, -1]] * 7, 1234.5, 4e-12, 3.0)
    a = O.aggregate(res, rows, SPEC, literal=True, ray_total=100)
    delay = 1234.5 / SPEC.cspeed
    assert (a["npath"] == 7).all() and (a["path_match"] == 0).all()
    assert np.allclose(a["results"]["power"], 4e-12, rtol=1e-14) and np.allclose(a["delay"], delay, rtol=1e-15)
    phase = -np.fmod(delay * 2 * np.pi * SPEC.carrier, 2 * np.pi)
    assert np.allclose(a["phase"], phase, rtol=1e-12) and np.allclose(a["results"]["doppler"], 3.0)
    assert list(O.unique_paths(a["path_match"])) == [0]


def test_two_paths_interleaved_and_two_receivers():
    rows = [[0, -1, -1], [1, -1, -1]] * 4
    res, rows = make(8, 0, rows, np.linspace(100, 800, 8), np.linspace(1e-10, 8e-10, 8))
    res["received"][6:] = 1
    a = O.aggregate(res, rows, SPEC, literal=True, ray_total=100)
    assert list(a["path_match"]) == [0, 1, 0, 1, 0, 1, 6, 7]
    assert list(a["npath"]) == [3, 3, 3, 3, 3, 3, 1, 1]
    assert list(O.unique_paths(a["path_match"])) == [0, 1, 6, 7]
    # mean voltage squared (aggregation.cu:89)
    v = np.sqrt(res["power"])
    assert np.isclose(a["results"]["power"][0], (v[[0, 2, 4]].mean()) ** 2, rtol=1e-14)


def test_direct_ray_aggregates_with_everything():
    """Appendix B-Q9 (aggregation.cu:56): a direct ray sums over every ray of its receiver."""
    res, rows = make(5, 0, [[-1, -1, -1], [0, -1, -1], [0, -1, -1], [1, 0, -1], [-1, -1, -1]], 100.0, 1e-10)
    res["reflDepth"][[0, 4]] = 0
    a = O.aggregate(res, rows, SPEC, literal=True, ray_total=50)
    assert list(a["npath"]) == [5, 2, 2, 1, 5]
    assert list(a["path_match"]) == [0, 1, 1, 3, 0]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_literal_equals_binned(seed):
    rng = np.random.default_rng(seed)
    R, D = 400, 3
    rows = rng.integers(-1, 3, size=(R, D)).astype(np.int32)
    res, _ = make(R, rng.integers(0, 3, size=R), rows, rng.uniform(10, 5000, R), rng.uniform(1e-14, 1e-9, R), rng.normal(0, 100, R),
                  refl=rng.integers(0, 3, size=R))
    a = O.aggregate(res, rows, SPEC, literal=True, ray_total=R)
    b = O.aggregate(res, rows, SPEC, literal=False, ray_total=R)
    assert np.array_equal(a["path_match"], b["path_match"]) and np.array_equal(a["npath"], b["npath"])
    for k in ("delay", "phase"):
        assert np.allclose(a[k], b[k], rtol=1e-12, atol=0)
    for k in ("power", "doppler"):
        assert np.allclose(a["results"][k], b["results"][k], rtol=1e-11, atol=1e-300)


def test_trace_bins_equals_postprocess_plus_literal():
    """The oracle's fused bins == trace -> host post-process (ray_tracer.cpp:1190-1258) -> literal aggregation."""
    from rts_b200 import scenes
    targets, spec = scenes.slab(n=24)
    r = O.trace(targets, spec)
    rx_res, rx_rows, rx_slots = O.postprocess(r["results"], r["targ_intersect"], spec)
    a = O.aggregate(rx_res, rx_rows, spec, literal=True)
    bins, _ = O.trace_bins(targets, spec, use_bvh=False)
    uniq = O.unique_paths(a["path_match"])
    assert len(uniq) == len(bins)
    for u in uniq:
        row = rx_rows[u]
        match = [b for b in bins if b["rx"] == rx_res["received"][u] and list(b["path"][:spec.depth_total]) == list(row)]
        assert len(match) == 1
        b = match[0]
        assert b["npath"] == a["npath"][u] and b["min_slot"] == rx_slots[u]
        assert np.isclose(b["power"], a["results"]["power"][u], rtol=1e-12) and np.isclose(b["delay"], a["delay"][u], rtol=1e-13)
        assert np.isclose(b["phase"], a["phase"][u], rtol=1e-11) and np.isclose(b["doppler"], a["results"]["doppler"][u], rtol=1e-11)


@pytest.mark.parametrize("scene", ["slab", "direct+", "direct-"])
def test_responses_with_rcs_and_gains(scene):
    """Fused bins with per-target RCS and Gt/Gr -> responses == trace -> post-process (ray_tracer.cpp:1190-1258)
    -> literal aggregation -> sort+unique(pathMatch) -> InterpPoint arguments (ray_tracer.cpp:1289-1320)."""
    from rts_b200 import scenes
    if scene == "slab":
        targets, spec = scenes.slab(n=24)
    else:
        targets, spec = scenes.direct_and_plate(n=96, side=1 if scene == "direct+" else -1)
    spec.targ_rcs = np.array([2.5 + 0.75 * k for k in range(len(targets))])
    spec.gain_tx, spec.gain_rx = 31.0, 7.5
    r = O.trace(targets, spec)
    rx_res, rx_rows, rx_slots = O.postprocess(r["results"], r["targ_intersect"], spec, rcs_per_target=spec.targ_rcs,
                                              gain=spec.gain_tx * spec.gain_rx)
    assert len(rx_res) > 0
    a = O.aggregate(rx_res, rx_rows, spec, literal=True)
    want = O.responses(a, rx_slots)
    bins, _ = O.trace_bins(targets, spec, use_bvh=False)
    got = O.responses_from_bins(bins)
    assert len(got) == len(want) > 0
    if scene.startswith("direct"):
        direct = [b for b in bins if b["direct"]]
        assert len(direct) == 1 and len(bins) == 2
        assert (len(got) == 2) == (scene == "direct+")   # the direct bin emits its own response only when a direct ray comes first
    assert np.array_equal(got["rx"], want["rx"]) and np.array_equal(got["slot"], want["slot"])
    for f in ("power", "delay", "doppler", "phase"):
        assert np.allclose(got[f], want[f], rtol=1e-11, atol=1e-300), f
