"""The rs::RTS adapter (include/rts_soars_adapter.hpp) against a stand-in host simulator (tests/mock_soars/).

CPU: the adapter compiles against a SOARS-shaped interface and links librts_b200.so.
GPU: the responses it pushes into the receivers equal the reference's host flow evaluated with the oracle:
trace -> per-ray callbacks (Target::GetRCS, GetGain; ray_tracer.cpp:1190-1258) -> literal aggregation
(aggregation.cu:32-97) -> unique paths -> InterpPoint arguments (ray_tracer.cpp:1289-1320)."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

import oracle_api as O
from rts_b200 import lib as L
from rts_b200.abi import PulseSpec, Target

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock_soars")
EXE = os.path.join(MOCK, "mock_main")


def build_mock():
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(MOCK, "mock_main.cpp"), "-o", EXE,
           "-L" + os.path.join(ROOT, "rts_b200"), "-lrts_b200", "-Wl,-rpath," + os.path.join(ROOT, "rts_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_adapter_compiles_and_links_against_a_soars_shaped_interface():
    exe = build_mock()
    assert os.path.exists(exe)


# ---- the mock world of tests/mock_soars/mock_main.cpp, mirrored ----
C0, CARRIER, TEMP = 299792458.0, 10e9, 290.0
SAMPLE = 1.0 / 1000.0
PRI = 2e-3
TX = dict(pos=np.zeros(3), az0=0.02, el0=0.01, az_rate=2.0, span=(0.5, 0.3, 0.0), peak=30.0)
RXS = [dict(pos=np.array([0.0, 0, 0]), az=0.0, el=0.0, sphere=(25.0, 2.5, 2.5), noise=120.0, peak=12.0),
       dict(pos=np.array([10.0, -60, 5]), az=1.2, el=0.0, sphere=(30.0, 3.0, 3.0), noise=80.0, peak=12.0)]
TARGETS = [dict(shape="rect", dims=(1.0, 30.0, 20.0), pos0=(100.0, 0, 0), vel=(0.0, 0, 0), rot0=(0.05, 0.0, 0.0), rate=(0.0, 0, 0), rotating=False, refl=0.8, refr=1.0, rcs0=3.0),
           dict(shape="sphere", subdivs=2, radius=4.0, pos0=(60.0, 12, 3), vel=(-40.0, 25, 10), rot0=(0.0, 0, 0), rate=(0.0, 0, 0), rotating=False, refl=0.9, refr=1.0, rcs0=1.5),
           dict(shape="rect", dims=(6.0, 4.0, 3.0), pos0=(70.0, -14, -2), vel=(15.0, 30, -5), rot0=(0.3, 0.1, -0.2), rate=(40.0, -15.0, 25.0), rotating=True, refl=0.7, refr=1.6, rcs0=0.8)]


def svec(v):
    n = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return (math.atan2(v[1], v[0]), math.asin(v[2] / n)) if n else (0.0, 0.0)


def pattern(angle, ref, peak):
    return peak * (1.0 + 0.25 * math.cos(angle[0] - ref[0]) * math.cos(angle[1] - ref[1]))


def world_at(t):
    out, vel = [], []
    for T in TARGETS:
        f = np.float32
        if T["shape"] == "rect":
            v, tri, nrm = L.rect_mesh(*T["dims"], yaw=f(T["rot0"][0]), pitch=f(T["rot0"][1]), roll=f(T["rot0"][2]))
        else:
            v, tri, nrm = L.sphere_mesh(T["subdivs"], T["radius"], yaw=f(T["rot0"][0]), pitch=f(T["rot0"][1]), roll=f(T["rot0"][2]))
        if T["rotating"] and t > 0.0:
            ang = [T["rot0"][a] + T["rate"][a] * t for a in range(3)]
            R = L.rotation_matrix(f(ang[0]), f(ang[1]), f(ang[2]))
            v = np.stack([((0.0 + R[i, 0] * v[:, 0]) + R[i, 1] * v[:, 1]) + R[i, 2] * v[:, 2] for i in range(3)], axis=1)
            nrm = np.stack([((0.0 + R[i, 0] * nrm[:, 0]) + R[i, 1] * nrm[:, 1]) + R[i, 2] * nrm[:, 2] for i in range(3)], axis=1)
        p0 = np.array([T["pos0"][a] + T["vel"][a] * t for a in range(3)])
        p1 = np.array([T["pos0"][a] + T["vel"][a] * (t + SAMPLE) for a in range(3)])
        out.append(Target(v + p0[None, :], tri, nrm, T["refl"], T["refr"]))
        vel.append((p1 - p0) / SAMPLE)
    return out, np.array(vel)


def expected_responses(N, max_refl, max_refr, pulses):
    per_rx = [[] for _ in RXS]
    Wl = C0 / CARRIER
    for k in range(pulses):
        t = k * PRI
        targets, vel = world_at(t)
        tx_rot = (TX["az0"] + TX["az_rate"] * t, TX["el0"])
        rx = [L.rx_sphere_from_desc(r["pos"], r["az"], r["el"], *r["sphere"]) for r in RXS]
        spec = PulseSpec(grid=(N, N, N), max_refl=max_refl, max_refr=max_refr, interpolate_smooth=False, tx_origin=tuple(TX["pos"]),
                         tx_dir=tx_rot, tx_span=TX["span"], cspeed=C0, carrier=CARRIER, rx=rx, targ_vel=vel)
        r = O.trace(targets, spec)
        res, ti, ang = r["results"], r["targ_intersect"], r["rcs_angle"]
        keep = np.nonzero(res["received"] >= 0)[0]
        if not len(keep):
            continue
        rx_res = res[keep].copy()
        rows = np.ascontiguousarray(ti[keep])
        for i, s in enumerate(keep):
            rec = rx_res[i]
            repos = RXS[rec["received"]]["pos"]
            if rec["reflDepth"] == 0 and rec["refrDepth"] == 0:
                tv, rv = TX["pos"] - repos, repos - TX["pos"]
            else:
                tv, rv = rec["firstHitPoint"] - TX["pos"], rec["prevHitPoint"] - repos
            p = rec["power"]
            for c in range(spec.depth_total):
                if rows[i, c] >= 0:
                    p *= TARGETS[rows[i, c]]["rcs0"] * (2.0 + 0.5 * math.cos(ang[s, c, 0]) + 0.25 * math.sin(ang[s, c, 1]))
            Gt = pattern(svec(tv), tx_rot, TX["peak"])
            Gr = pattern(svec(rv), (RXS[rec["received"]]["az"], RXS[rec["received"]]["el"]), RXS[rec["received"]]["peak"])
            p *= (Wl * Wl * Gt * Gr)
            Vr = rec["doppler"] / 2
            rx_res[i]["power"] = p
            rx_res[i]["doppler"] = CARRIER * (((1 + Vr / C0) / (1 - Vr / C0)) - 1)
        a = O.aggregate(rx_res, rows, spec, literal=True)
        for resp in O.responses(a, keep.astype(np.uint64)):
            j = int(resp["rx"])
            per_rx[j].append(dict(rx=j, power=float(resp["power"]), time=t + float(resp["delay"]), delay=float(resp["delay"]),
                                  doppler=float(resp["doppler"]), phase=float(resp["phase"]), noise=TEMP + RXS[j]["noise"]))
    return [x for lst in per_rx for x in lst]


@pytest.mark.gpu
@pytest.mark.parametrize("args", [(24, 2, 0, 3), (20, 2, 2, 2)])
def test_adapter_responses_equal_the_reference_host_flow(args):
    exe = build_mock()
    N, max_refl, max_refr, pulses = args
    r = subprocess.run([exe, "exact"] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    want = expected_responses(N, max_refl, max_refr, pulses)
    assert len(got) == len(want) > 0
    for g, w in zip(got, want):
        assert g["rx"] == w["rx"] and g["noise"] == w["noise"]
        for f in ("power", "time", "delay", "phase"):
            assert math.isclose(g[f], w[f], rel_tol=1e-9), (f, g, w)
        assert math.isclose(g["doppler"], w["doppler"], rel_tol=1e-9, abs_tol=1e-9), (g, w)


@pytest.mark.gpu
def test_adapter_fused_mode_runs_and_matches_exact_paths_delays():
    """Fused mode folds scalar RCS / boresight gains on the device: same unique paths, delays and phases as the exact
    path (powers differ by the angle dependence of the mock callbacks)."""
    exe = build_mock()
    a = subprocess.run([exe, "exact", "24", "2", "0", "2"], capture_output=True, text=True, timeout=600)
    b = subprocess.run([exe, "fused", "24", "2", "0", "2"], capture_output=True, text=True, timeout=600)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    ga = [json.loads(l) for l in a.stdout.splitlines() if l.startswith("{")]
    gb = [json.loads(l) for l in b.stdout.splitlines() if l.startswith("{")]
    assert len(ga) == len(gb) > 0
    for x, y in zip(ga, gb):
        assert x["rx"] == y["rx"] and math.isclose(x["delay"], y["delay"], rel_tol=1e-12) and math.isclose(x["phase"], y["phase"], rel_tol=1e-9)
        assert y["power"] > 0


@pytest.mark.gpu
def test_adapter_tabulated_mode_matches_the_exact_path():
    """Options::tabulated: the mock's angle-dependent GetRCS / GetGain sampled on 1441 x 721 grids and evaluated on the device
    (RTS_TABLES) — no per-ray data on the host — against the exact path that calls them per received ray.  The mock's
    callbacks are smooth cosines, so the difference is the bilinear interpolation error: h^2/8 with h = 4pi/1440 -> 1e-5."""
    exe = build_mock()
    a = subprocess.run([exe, "exact", "24", "2", "0", "3"], capture_output=True, text=True, timeout=600)
    b = subprocess.run([exe, "tabulated", "24", "2", "0", "3"], capture_output=True, text=True, timeout=600)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    ga = [json.loads(l) for l in a.stdout.splitlines() if l.startswith("{")]
    gb = [json.loads(l) for l in b.stdout.splitlines() if l.startswith("{")]
    assert len(ga) == len(gb) > 0
    for x, y in zip(ga, gb):
        assert x["rx"] == y["rx"] and x["noise"] == y["noise"]
        assert math.isclose(x["delay"], y["delay"], rel_tol=1e-12) and math.isclose(x["phase"], y["phase"], rel_tol=1e-9)
        assert math.isclose(x["doppler"], y["doppler"], rel_tol=1e-9, abs_tol=1e-9)
        assert math.isclose(x["power"], y["power"], rel_tol=5e-5), (x, y)
