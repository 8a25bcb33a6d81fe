"""Boundary behaviour of the C-ABI on a GPU box: errors come back as status codes with a message —
the reference aborts instead (aggregation.cu:17-27, ray_tracer.cpp:455-458) — distinct handles work from distinct host
threads, and the measurement aid behind bench.py's roofline.l2 returns sane numbers."""
import ctypes as C
import threading

import numpy as np
import pytest

import oracle_api as O
import parity
from rts_b200 import abi, lib as L
from rts_b200 import scenes

pytestmark = pytest.mark.gpu

RTS_ERR_ARG, RTS_ERR_STATE, RTS_ERR_CAPACITY = -1, -4, -5      # include/rts_b200.h


def test_errors_are_status_codes_not_exits(engine):
    l = L.load()
    eng = L.Engine(0)
    try:
        # results asked for before any pulse
        with pytest.raises(L.RtsError, match="no pulse|did not produce"):
            eng.bins()
        with pytest.raises(L.RtsError, match="no pulse"):
            eng.wave_profile()
        # pulse whose velocity array does not match the committed scene
        t, s = scenes.flat_plate(n=16)
        eng.set_targets(t)
        cp = abi.CPulse(s, len(t))
        cp.c.n_targets = len(t) + 2
        assert l.rts_trace_pulse(eng._h, C.byref(cp.c), L.RTS_OUT_BINS) == RTS_ERR_ARG
        assert b"target" in l.rts_last_error()
        cp.c.n_targets = len(t)
        assert l.rts_trace_pulse(eng._h, C.byref(cp.c), 0) == RTS_ERR_ARG            # no output requested
        cp.c.max_refl = 40
        assert l.rts_trace_pulse(eng._h, C.byref(cp.c), L.RTS_OUT_BINS) == RTS_ERR_CAPACITY
        # poses for the wrong number of targets
        with pytest.raises(L.RtsError):
            eng.set_poses([None] * (len(t) + 1), [(0, 0, 0)] * (len(t) + 1))
        # NULL handle / NULL outputs
        assert l.rts_sync(None) == RTS_ERR_ARG
        assert l.rts_get_bins(None, None, 0, None) == RTS_ERR_ARG
        assert l.rts_kernel_launches(eng._h, None) == RTS_ERR_ARG
        assert l.rts_probe_read_bandwidth(eng._h, 16, 1, None) == RTS_ERR_ARG
        assert l.rts_last_error()                              # thread-local message is set
        # the handle is still usable after every failure
        st = eng.trace(s, L.RTS_OUT_BINS)
        assert st["primary_rays"] == 16 * 16 and len(eng.bins()) >= 1
        # bins only exist for a pulse traced with RTS_OUT_BINS
        eng.trace(s, L.RTS_OUT_RECORDS)
        with pytest.raises(L.RtsError, match="did not produce bins"):
            eng.bins()
    finally:
        eng.close()


def test_two_engines_on_two_host_threads():
    """One handle per thread (the reference is one context per call, not re-entrant: SURVEY §8b 'Threading')."""
    cases = {"plate": scenes.flat_plate(n=128, max_refl=3), "trihedral": scenes.trihedral(n=160)}
    want = {k: O.trace_bins(t, s)[0] for k, (t, s) in cases.items()}
    got, errors = {}, []

    def work(name):
        try:
            t, s = cases[name]
            with L.Engine(0) as eng:
                eng.set_targets(t)
                for _ in range(6):                      # interleave with the other thread's launches
                    eng.trace(s, L.RTS_OUT_BINS)
                got[name] = eng.bins()
        except Exception as ex:                         # noqa: BLE001 - reported below
            errors.append((name, repr(ex)))

    threads = [threading.Thread(target=work, args=(k,)) for k in cases]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for k in cases:
        parity.assert_bins_close(parity.compare_bins(got[k], want[k]))


def test_read_bandwidth_probe(engine):
    l2 = engine.probe_read_bandwidth(32 << 20, 50)
    hbm = engine.probe_read_bandwidth(1 << 30, 3)
    assert 3000.0 < hbm < 9000.0, hbm            # B200 HBM3e: ~6.5-7.5 TB/s achievable
    assert l2 > 1.5 * hbm, (l2, hbm)             # an L2-resident buffer streams well above HBM rate
