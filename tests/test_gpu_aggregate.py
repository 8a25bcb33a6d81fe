"""The drop-in aggregator (rts_aggregate / rs::kernel_wrapper, replacing aggregation.cu) against the
literal transcription in the oracle and, when oracle/_ref/libref_aggregation.so was built, against the
reference's own aggregation.cu compiled unmodified for sm_100a."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_api as O
import parity
from rts_b200 import scenes
from rts_b200.abi import RAY_RECORD, PulseSpec

pytestmark = pytest.mark.gpu


def _random_case(seed, R=3000, D=4, n_rx=3, n_targ=3):
    rng = np.random.default_rng(seed)
    res = np.zeros(R, dtype=RAY_RECORD)
    res["received"] = rng.integers(0, n_rx, size=R)
    res["rayLength"] = rng.uniform(10, 9000, R)
    res["power"] = rng.uniform(1e-16, 1e-9, R)
    res["doppler"] = rng.normal(0, 300, R)
    res["reflDepth"] = rng.integers(0, 3, size=R)
    rows = rng.integers(-1, n_targ, size=(R, D)).astype(np.int32)
    return res, rows


def _check(a, b, counts_exact=True):
    assert np.array_equal(a["path_match"], b["path_match"])
    assert np.array_equal(a["npath"], b["npath"])
    for k in ("delay", "phase"):
        assert np.allclose(a[k], b[k], rtol=1e-5, atol=0)
    for k in ("power", "doppler"):
        assert np.allclose(a["results"][k], b["results"][k], rtol=1e-5, atol=1e-300)


@pytest.mark.parametrize("seed", [0, 1])
def test_aggregate_matches_literal_oracle(engine, seed):
    res, rows = _random_case(seed)
    spec = PulseSpec(grid=(1, 1, 1), max_refl=2, max_refr=2)
    a = engine.aggregate(res, rows, spec.cspeed, spec.carrier, ray_total=len(res))
    b = O.aggregate(res, rows, spec, literal=True, ray_total=len(res))
    _check(a, b)
    assert list(O.unique_paths(a["path_match"])) == list(O.unique_paths(b["path_match"]))


def test_aggregate_edge_cases(engine):
    spec = PulseSpec(grid=(1, 1, 1), max_refl=3, max_refr=0)
    # one ray; all identical; a direct ray among reflected ones (Appendix B-Q9); D = 0
    for res, rows in [
        _random_case(5, R=1, D=3),
        (np.repeat(_random_case(6, R=1, D=3)[0], 50), np.repeat(_random_case(6, R=1, D=3)[1], 50, axis=0)),
        _random_case(7, R=257, D=1, n_rx=1, n_targ=1),
    ]:
        a = engine.aggregate(res, rows, spec.cspeed, spec.carrier, ray_total=1000)
        b = O.aggregate(res, rows, spec, literal=True, ray_total=1000)
        _check(a, b)
    res, _ = _random_case(8, R=40)
    res["reflDepth"] = 0
    rows0 = np.zeros((40, 0), dtype=np.int32)
    a = engine.aggregate(res, rows0, spec.cspeed, spec.carrier, ray_total=1000)
    b = O.aggregate(res, rows0, spec, literal=True, ray_total=1000)
    _check(a, b)
    # empty input is a no-op
    engine.aggregate(res[:0], rows0[:0], spec.cspeed, spec.carrier, ray_total=10)


def test_two_phase_path_equals_fused_bins(engine):
    """trace (records) -> host post-process (ray_tracer.cpp:1190-1258) -> aggregate == fused bins."""
    from rts_b200 import lib as L
    targets, spec = scenes.slab(n=80)
    engine.set_targets(targets)
    engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
    res, ti, _, _ = engine.records(rcs=False, tri_path=False)
    bins = engine.bins()
    rx_res, rx_rows, rx_slots = O.postprocess(res, ti, spec)
    a = engine.aggregate(rx_res, rx_rows, spec.cspeed, spec.carrier, ray_total=spec.ray_total)
    uniq = O.unique_paths(a["path_match"])
    assert len(uniq) == len(bins)
    for u in uniq:
        m = [b for b in bins if b["rx"] == rx_res["received"][u] and list(b["path"][:spec.depth_total]) == list(rx_rows[u])]
        assert len(m) == 1 and m[0]["npath"] == a["npath"][u] and m[0]["min_slot"] == rx_slots[u]
        assert np.isclose(m[0]["power"], a["results"]["power"][u], rtol=1e-10) and np.isclose(m[0]["phase"], a["phase"][u], rtol=1e-10)


@pytest.mark.skipif(not os.path.exists(O.REF_AGG_LIB), reason="oracle/_ref/libref_aggregation.so not built")
@pytest.mark.parametrize("seed", [11, 12])
def test_aggregate_matches_reference_aggregation_cu(engine, seed):
    """Second opinion from the reference's own aggregation.cu (compiled unmodified, run on this GPU)."""
    ref = C.CDLL(O.REF_AGG_LIB)
    res, rows = _random_case(seed, R=2000)
    spec = PulseSpec(grid=(1, 1, 1), max_refl=2, max_refr=2)
    R, D = len(res), rows.shape[1]
    r2 = res.copy()
    acc = {k: np.zeros(R) for k in ("npath", "power", "doppler", "delay", "phase")}
    pm = np.full(R, R + 1, dtype=np.int32)
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    ref.ref_kernel_wrapper.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_double] + \
        [C.POINTER(C.c_double)] * 5 + [C.POINTER(C.c_int)]
    ref.ref_kernel_wrapper(r2.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.POINTER(C.c_int)), R, D, 256, 1024, spec.cspeed, spec.carrier,
                           dp(acc["npath"]), dp(acc["power"]), dp(acc["doppler"]), dp(acc["delay"]), dp(acc["phase"]),
                           pm.ctypes.data_as(C.POINTER(C.c_int)))
    a = engine.aggregate(res, rows, spec.cspeed, spec.carrier, ray_total=R)
    # the reference copies back results, delay, phase and pathMatch only (aggregation.cu:169-172)
    assert np.array_equal(a["path_match"], pm)
    assert np.allclose(a["delay"], acc["delay"], rtol=1e-5) and np.allclose(a["phase"], acc["phase"], rtol=1e-5)
    assert np.allclose(a["results"]["power"], r2["power"], rtol=1e-5) and np.allclose(a["results"]["doppler"], r2["doppler"], rtol=1e-5, atol=1e-9)
    # and the oracle's literal transcription agrees with the reference kernel too
    b = O.aggregate(res, rows, spec, literal=True, ray_total=R)
    assert np.array_equal(b["path_match"], pm) and np.allclose(b["delay"], acc["delay"], rtol=1e-12)


def test_kernel_wrapper_by_its_reference_symbol(engine):
    """rs::kernel_wrapper called the way a host linked against the reference calls it — through the C++ symbol with the
    reference's signature (aggregation.cuh:18-23), host arrays in and out, no engine handle — from two threads (each gets
    its own engine, released when the thread ends) and compared with rts_aggregate and the literal oracle."""
    import threading
    from rts_b200 import lib as L
    lib = L.load()
    fn = getattr(lib, "_ZN2rs14kernel_wrapperEP10PerRayDataPijjjjddPdS3_S3_S3_S3_S2_")
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_double] + \
        [C.POINTER(C.c_double)] * 5 + [C.POINTER(C.c_int)]
    spec = PulseSpec(grid=(1, 1, 1), max_refl=2, max_refr=2)
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    out = {}

    def call(seed):
        res, rows = _random_case(seed, R=2500)
        R, D = len(res), rows.shape[1]
        for _ in range(2):                                         # the second call reuses the thread's engine
            r2 = res.copy()                                        # fresh arrays per call, as the reference's host loop has them
            acc = {k: np.zeros(R) for k in ("npath", "power", "doppler", "delay", "phase")}
            pm = np.full(R, R + 1, dtype=np.int32)                 # ray_tracer.cpp:1271
            fn(r2.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.POINTER(C.c_int)), R, D, 256, 1024, spec.cspeed, spec.carrier,
               dp(acc["npath"]), dp(acc["power"]), dp(acc["doppler"]), dp(acc["delay"]), dp(acc["phase"]), pm.ctypes.data_as(C.POINTER(C.c_int)))
        out[seed] = (res, rows, r2, acc, pm)

    threads = [threading.Thread(target=call, args=(s,)) for s in (21, 22)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for seed in (21, 22):
        res, rows, r2, acc, pm = out[seed]
        a = engine.aggregate(res, rows, spec.cspeed, spec.carrier, ray_total=len(res))
        b = O.aggregate(res, rows, spec, literal=True, ray_total=len(res))
        for want in (a, b):
            assert np.array_equal(want["path_match"], pm)
            assert np.allclose(want["delay"], acc["delay"], rtol=1e-5) and np.allclose(want["phase"], acc["phase"], rtol=1e-5)
            assert np.allclose(want["results"]["power"], r2["power"], rtol=1e-5, atol=1e-300)
            assert np.allclose(want["results"]["doppler"], r2["doppler"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("scene", ["slab", "direct+", "direct-"])
def test_fused_rcs_gains_and_responses(engine, scene):
    """RTS_OUT_BINS with per-target RCS and Gt/Gr (the SOARS callbacks of ray_tracer.cpp:1219-1247 as scalars) and
    rts_get_responses == oracle trace -> host post-process -> literal aggregation -> sort+unique -> InterpPoint
    arguments (ray_tracer.cpp:1289-1320)."""
    from rts_b200 import lib as L
    if scene == "slab":
        targets, spec = scenes.slab(n=64)
    else:
        targets, spec = scenes.direct_and_plate(n=96, side=1 if scene == "direct+" else -1)
    spec.targ_rcs = np.array([2.5 + 0.75 * k for k in range(len(targets))])
    spec.gain_tx, spec.gain_rx = 31.0, 7.5
    engine.set_targets(targets)
    engine.trace(spec, L.RTS_OUT_BINS)
    gbins, got = engine.bins(), engine.responses()
    obins, _ = O.trace_bins(targets, spec, use_bvh=False)
    parity.assert_bins_close(parity.compare_bins(gbins, obins))
    assert np.array_equal(gbins["own_min_slot"], obins["own_min_slot"])
    r = O.trace(targets, spec)
    rx_res, rx_rows, rx_slots = O.postprocess(r["results"], r["targ_intersect"], spec, rcs_per_target=spec.targ_rcs,
                                              gain=spec.gain_tx * spec.gain_rx)
    a = O.aggregate(rx_res, rx_rows, spec, literal=True)
    want = O.responses(a, rx_slots)
    assert len(got) == len(want) > 0
    assert np.array_equal(got["rx"], want["rx"]) and np.array_equal(got["slot"], want["slot"])
    for f in ("power", "delay", "doppler", "phase"):
        assert np.allclose(got[f], want[f], rtol=1e-5, atol=1e-300), f


@pytest.mark.parametrize("scene", ["slab", "direct+", "empty"])
def test_received_rays_compacted_on_the_device(engine, scene):
    """rts_get_received == the received slots of the full reference-shaped arrays, in slot order
    (ray_tracer.cpp:1190-1221), and feeding them through the callbacks + rts_aggregate reproduces the fused bins."""
    from rts_b200 import lib as L
    if scene == "slab":
        targets, spec = scenes.slab(n=64)
    elif scene == "direct+":
        targets, spec = scenes.direct_and_plate(n=96, side=1)
    else:
        targets, spec = scenes.flat_plate(n=32)
        spec.rx = [L.rx_sphere_from_desc((0.0, 500.0, 0.0), 0.0, 0.0, 1.0, 0.1, 0.1)]   # nothing is received
    engine.set_targets(targets)
    engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_OUT_BINS)
    res, ti, rcs, _ = engine.records(tri_path=False)
    got_res, got_ti, got_rcs, got_slots = engine.received()
    keep = np.nonzero(res["received"] >= 0)[0]
    assert np.array_equal(got_slots, keep.astype(np.uint64))
    for f in res.dtype.names:                                        # field by field: numpy does not copy padding bytes
        assert got_res[f].tobytes() == res[keep][f].tobytes(), f
    assert np.array_equal(got_ti, ti[keep]) and got_rcs.tobytes() == rcs[keep].tobytes()
    if len(keep):
        rx_res, rx_rows, rx_slots = O.postprocess(res, ti, spec)      # RCS = gains = 1 callbacks + Doppler conversion
        a = engine.aggregate(rx_res, rx_rows, spec.cspeed, spec.carrier, ray_total=spec.ray_total)
        want = O.responses(a, rx_slots)
        got = engine.responses()
        assert np.array_equal(got["slot"], want["slot"]) and np.allclose(got["power"], want["power"], rtol=1e-10)
