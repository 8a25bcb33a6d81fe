"""SURVEY.md §8 f-2 on the device: angle-dependent Target::GetRCS and Transmitter/Receiver::GetGain as tables
(rts_set_rcs_tables / rts_set_antennas, RTS_TABLES) folded into the fused bins, against the two-phase path — records from
the GPU, the reference's per-ray host loop (ray_tracer.cpp:1190-1258) restated in numpy with the SAME tables as its
callbacks, rts_aggregate — to 1e-5 relative (the bar of BASELINE.json for sums)."""
import math

import numpy as np
import pytest

import oracle_api as O
from rts_b200 import lib as L, scenes
from rts_b200.abi import Antenna, Table2d

pytestmark = pytest.mark.gpu


def _rcs_table(k, n_az=73, n_el=37):
    az = np.linspace(-2 * math.pi, 2 * math.pi, n_az)
    el = np.linspace(-math.pi, math.pi, n_el)
    A, E = np.meshgrid(az, el, indexing="ij")
    vals = (1.5 + 0.4 * k) * (2.0 + 0.8 * np.cos(A * (1 + 0.3 * k)) + 0.5 * np.sin(E + 0.2 * k))
    return Table2d(az[0], az[1] - az[0], el[0], el[1] - el[0], vals)


def _gain_table(peak, n=61):
    d = np.linspace(-math.pi, math.pi, n)
    A, E = np.meshgrid(d, d, indexing="ij")
    return Table2d(d[0], d[1] - d[0], d[0], d[1] - d[0], peak * (1.0 + 0.25 * np.cos(A) * np.cos(E)) * np.exp(-0.05 * (A * A + E * E)))


def _svec(v):
    ln = np.sqrt((v * v).sum(axis=1))
    ok = ln != 0
    az = np.where(ok, np.arctan2(v[:, 1], v[:, 0]), 0.0)
    el = np.where(ok, np.arcsin(np.where(ok, v[:, 2] / np.where(ok, ln, 1.0), 0.0)), 0.0)
    return az, el


def host_postprocess(res, ti, ang, spec, rcs_fns, tx_ant, rx_ants, rcs_scalar=None):
    """ray_tracer.cpp:1190-1258 over the received rays, the callbacks being the tables (or scalars)."""
    keep = np.nonzero(res["received"] >= 0)[0]
    r = res[keep].copy()
    rows, a = ti[keep], ang[keep]
    Wl = spec.cspeed / spec.carrier
    power = r["power"].copy()
    for k in range(rows.shape[1]):                                  # :1221-1231, column order
        for t in range(len(rcs_fns)):
            m = rows[:, k] == t
            if m.any():
                power[m] *= rcs_fns[t](a[m, k, 0], a[m, k, 1]) if rcs_fns[t] is not None else (rcs_scalar[t] if rcs_scalar is not None else 1.0)
    origin = np.asarray(spec.tx_origin, dtype=np.float64)
    repos = np.array([rx_ants[j].position for j in r["received"]], dtype=np.float64).reshape(-1, 3)
    direct = (r["reflDepth"] == 0) & (r["refrDepth"] == 0)
    tv = np.where(direct[:, None], origin[None, :] - repos, r["firstHitPoint"] - origin[None, :])            # :1205-1212
    rv = np.where(direct[:, None], repos - origin[None, :], r["prevHitPoint"] - repos)
    delay = r["rayLength"] / spec.cspeed
    taz, tel = _svec(tv)
    raz, rel_ = _svec(rv)
    Gt = tx_ant.gain(taz - tx_ant.bore_az, tel - tx_ant.bore_el) if tx_ant is not None and tx_ant.gain is not None else np.full(len(r), spec.gain_tx or 1.0)
    Gr = np.empty(len(r))
    for j, ant in enumerate(rx_ants):
        m = r["received"] == j
        if ant.gain is not None:
            Gr[m] = ant.gain(raz[m] - (ant.bore_az + ant.rate_az * delay[m]), rel_[m] - (ant.bore_el + ant.rate_el * delay[m]))
        else:
            Gr[m] = spec.gain_rx or 1.0
    power *= (Wl * Wl * Gt * Gr)                                    # :1247
    Vr = r["doppler"] / 2
    r["doppler"] = spec.carrier * (((1 + Vr / spec.cspeed) / (1 - Vr / spec.cspeed)) - 1)   # :1252-1253
    r["power"] = power
    return r, rows, keep.astype(np.uint64)


def _compare(engine, bins, rx_res, rx_rows, rx_slots, spec):
    a = engine.aggregate(rx_res, rx_rows, spec.cspeed, spec.carrier, ray_total=spec.ray_total)
    uniq = O.unique_paths(a["path_match"])
    assert len(uniq) == len(bins) > 0
    worst = 0.0
    for u in uniq:
        m = [b for b in bins if b["rx"] == rx_res["received"][u] and list(b["path"][:spec.depth_total]) == list(rx_rows[u])]
        assert len(m) == 1
        b = m[0]
        assert b["npath"] == a["npath"][u] and b["min_slot"] == rx_slots[u]
        for got, want in ((b["power"], a["results"]["power"][u]), (b["phase"], a["phase"][u]), (b["delay"], a["delay"][u]),
                          (b["doppler"], a["results"]["doppler"][u])):
            rel = abs(got - want) / max(abs(want), 1e-300)
            assert rel <= 1e-5 or abs(got - want) < 1e-12, (got, want)
            worst = max(worst, rel if abs(want) > 1e-9 else 0.0)
    return worst


def _antennas(spec, rx_positions, with_tx=True, with_rx=True):
    tx = Antenna(position=spec.tx_origin, bore_az=spec.tx_dir[0], bore_el=spec.tx_dir[1], gain=_gain_table(30.0)) if with_tx else None
    rx = [Antenna(position=tuple(p), bore_az=math.pi + 0.1 * j, bore_el=-0.05 * j, rate_az=40.0 * (j + 1), rate_el=-15.0,
                  gain=_gain_table(12.0 + j) if with_rx else None) for j, p in enumerate(rx_positions)]
    return tx, rx


@pytest.mark.parametrize("case", ["trihedral", "terrain", "direct"])
def test_tabulated_rcs_and_gains_match_the_two_phase_path(engine, case):
    if case == "trihedral":
        targets, spec = scenes.trihedral(n=160)
        rx_pos = [np.array(spec.tx_origin) + np.array([0.5, 0.3, 0.1])]
        poses = None
    elif case == "direct":
        targets, spec = scenes.direct_and_plate(n=96, side=1)
        rx_pos = [np.array(s.centre[:]) + np.array([0.2, -0.1, 0.3]) for s in spec.rx]
        poses = None
    else:
        ms = scenes.terrain_scene(n=384, cells_x=96, cells_y=48, movers=6, n_rx=3)
        targets, spec, poses = ms.base, ms.spec_for(2), ms.poses(2)
        rx_pos = [np.array(s.centre[:]) + np.array([30.0, -20.0, 10.0]) for s in spec.rx]
    rcs = [_rcs_table(k) if k % 3 != 2 else None for k in range(len(targets))]     # every third target keeps a scalar RCS
    scal = np.array([1.0 + 0.5 * k for k in range(len(targets))])
    spec.targ_rcs = scal
    tx, rx = _antennas(spec, rx_pos)
    engine.set_targets(targets)
    if poses is not None:
        engine.set_poses(*poses)
    try:
        engine.set_rcs_tables(rcs)
        engine.set_antennas(tx, rx)
        # two-phase: records -> host loop with the tables as callbacks -> rts_aggregate
        engine.trace(spec, L.RTS_OUT_RECORDS | L.RTS_NO_REUSE)
        res, ti, ang, _ = engine.records(tri_path=False)
        rx_res, rx_rows, rx_slots = host_postprocess(res, ti, ang, spec, rcs, tx, rx, rcs_scalar=scal)
        assert len(rx_res) > 50
        # fused, from scratch (k_primary_follow) and with between-pulse reuse (kept hits, k_wave1_kept)
        for flags in (L.RTS_NO_REUSE, 0, 0):
            engine.trace(spec, L.RTS_OUT_BINS | L.RTS_TABLES | flags)
            worst = _compare(engine, engine.bins(), rx_res, rx_rows, rx_slots, spec)
            assert worst <= 1e-9, worst               # observed: rounding only
        # the tables really are in use: without RTS_TABLES the powers differ
        engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_REUSE)
        plain = engine.bins()
        engine.trace(spec, L.RTS_OUT_BINS | L.RTS_TABLES | L.RTS_NO_REUSE)
        tab = engine.bins()
        assert len(plain) == len(tab) and not np.allclose(plain["power"], tab["power"], rtol=1e-3, atol=0)
        # RCS tables only / antennas only
        engine.set_antennas(None, None)
        engine.trace(spec, L.RTS_OUT_BINS | L.RTS_TABLES | L.RTS_NO_REUSE)
        r2, rows2, slots2 = host_postprocess(res, ti, ang, spec, rcs, None, [Antenna(position=tuple(p)) for p in rx_pos], rcs_scalar=scal)
        _compare(engine, engine.bins(), r2, rows2, slots2, spec)
        engine.set_rcs_tables(None)
        engine.set_antennas(*_antennas(spec, rx_pos, with_tx=False))
        spec2 = spec
        spec2.targ_rcs = None
        engine.trace(spec2, L.RTS_OUT_BINS | L.RTS_TABLES | L.RTS_NO_REUSE)
        _, rx_only = _antennas(spec, rx_pos, with_tx=False)
        r3, rows3, slots3 = host_postprocess(res, ti, ang, spec2, [None] * len(targets), None, rx_only)
        _compare(engine, engine.bins(), r3, rows3, slots3, spec2)
    finally:
        engine.set_rcs_tables(None)
        engine.set_antennas(None, None)


def test_tables_contract(engine):
    targets, spec = scenes.slab(n=32)              # refraction on
    engine.set_targets(targets)
    with pytest.raises(L.RtsError):                # nothing uploaded
        engine.trace(spec, L.RTS_OUT_BINS | L.RTS_TABLES)
    engine.set_rcs_tables([_rcs_table(k) for k in range(len(targets))])
    try:
        with pytest.raises(L.RtsError):            # tabulated RCS + refraction: the two-phase path is the exact one
            engine.trace(spec, L.RTS_OUT_BINS | L.RTS_TABLES)
        with pytest.raises(L.RtsError):            # records mode keeps the reference's unmodified power
            engine.trace(spec, L.RTS_OUT_BINS | L.RTS_OUT_RECORDS | L.RTS_TABLES)
        with pytest.raises(L.RtsError):            # wrong count
            engine.set_rcs_tables([_rcs_table(0)] * (len(targets) + 1))
            t2, s2 = scenes.trihedral(n=16)
            engine.trace(s2, L.RTS_OUT_BINS | L.RTS_TABLES)
        with pytest.raises(L.RtsError):            # a transmitter antenna alone
            engine.set_antennas(Antenna(position=(0, 0, 0), gain=_gain_table(3.0)), None)
        bad = _rcs_table(0)
        bad.az_step = 0.0
        with pytest.raises(L.RtsError):
            engine.set_rcs_tables([bad] * len(targets))
    finally:
        engine.set_rcs_tables(None)
        engine.set_antennas(None, None)
