"""Analytic known-answer tests of the oracle: the formulas are read off the
reference source, the expected numbers are derived independently with numpy."""
import math

import numpy as np
import pytest

import oracle_api as O
from rts_b200 import scenes
from rts_b200.abi import PulseSpec, Target

FOUR_PI = 4 * math.pi


def test_flat_plate_radar_equation():
    """C1: every ray hits the plate once; captured power = 1/((4 pi)^3 L1^2 R^2) (normal_shader.cu:164,
    ray_tracer.cu:416); rayLength = fl32(L1) + t (normal_shader.cu:153, ray_tracer.cu:417)."""
    targets, spec = scenes.flat_plate(n=64)
    r = O.trace(targets, spec)
    res = r["results"]
    assert r["stats"]["hits"] == spec.rays and (res["reflDepth"] == 1).all()
    assert (r["targ_intersect"][:, 0] == 0).all()
    P = res["firstHitPoint"]
    assert np.abs(P[:, 0] - 100.0).max() < 2e-5          # t is narrowed to fp32 (Appendix B-Q4)
    assert np.abs(P[:, 1]).max() < 4.6 and np.abs(P[:, 2]).max() < 4.6
    got = res[res["received"] >= 0]
    assert len(got) > 0 and (got["received"] == 0).all()
    L1 = np.linalg.norm(got["firstHitPoint"], axis=1)
    t = got["rayLength"] - L1.astype(np.float32).astype(np.float64)
    c = np.array(spec.rx[0].centre[:])
    # reflected direction is (-d.x, d.y, d.z): the end point lies on the receiver sphere
    d = got["firstHitPoint"] / L1[:, None]
    refl = d * np.array([-1.0, 1, 1])
    end = got["firstHitPoint"] + t[:, None] * refl
    assert np.abs(np.linalg.norm(end - c, axis=1) - 2.0).max() < 1e-4
    expect = 1.0 / (FOUR_PI ** 3 * L1 ** 2 * t ** 2)
    assert np.abs(got["power"] / expect - 1).max() < 1e-5
    # all captured rays share one (receiver, path) bin
    bins, _ = O.trace_bins(targets, spec, use_bvh=False)
    assert len(bins) == 1 and bins[0]["npath"] == len(got) and list(bins[0]["path"][:1]) == [0]


def test_trihedral_retro_reflection():
    """C2: interior rays bounce exactly three times, once per face, and come back anti-parallel."""
    targets, spec = scenes.trihedral(n=120)
    r = O.trace(targets, spec)
    res, rows = r["results"], r["targ_intersect"]
    three = res["reflDepth"] == 3
    assert three.mean() > 0.9
    assert (np.sort(rows[three], axis=1) == np.array([0, 1, 2])).all()
    got = res[(res["received"] >= 0)]
    assert len(got) > 0 and (got["reflDepth"] == 3).all()
    # total path to the plane through the Tx normal to the boresight equals 2 * (A . b)
    b = np.array([math.cos(spec.tx_dir[0]) * math.cos(spec.tx_dir[1]), math.sin(spec.tx_dir[0]) * math.cos(spec.tx_dir[1]), math.sin(spec.tx_dir[1])])
    A = np.array([60.0, 60.0, 60.0])
    first = got["firstHitPoint"]
    d_in = first / np.linalg.norm(first, axis=1)[:, None]
    # path length up to the last hit + distance back to the plane along -d_in
    last = got["prevHitPoint"]
    # length recorded includes the capture segment t; remove it using the geometry: end = last + t*(-d_in)
    # so project: remaining distance from `last` to the plane {x.b = 0} along -d_in
    s = (last @ b) / (d_in @ b)
    path_to_last = np.linalg.norm(first, axis=1)
    # reconstruct the two interior segments from the unfolded-image identity: total = 2 (A.b)/(d.b)
    total = 2 * (A @ b) / (d_in @ b)
    rows_got = r["targ_intersect"][res["received"] >= 0]
    assert rows_got.shape[1] == 3
    # verify with the recorded length: rayLength = (segments to last hit) + t, and end point on the Rx sphere
    c = np.array(spec.rx[0].centre[:])
    seg_sum = total - s                           # segments up to the last hit
    t_cap = got["rayLength"] - seg_sum
    end = last - t_cap[:, None] * d_in
    assert np.abs(np.linalg.norm(end - c, axis=1) - 3.0).max() < 5e-3


def test_no_fourth_hit_and_absorb_rule():
    """normal_shader.cu:134: with maxRefl=1 a second hit is absorbed with no state change."""
    targets, spec = scenes.trihedral(n=60, max_refl=1)
    r = O.trace(targets, spec)
    res = r["results"]
    assert (res["reflDepth"] <= 1).all()
    assert r["stats"]["hits"] > r["stats"]["shaded_hits"]          # absorbed second hits exist
    absorbed = r["tri_path"][:, 1] >= 0
    assert absorbed.any() and (res["received"][absorbed] == -1).all()
    # absorbed rays keep the state of their first bounce: prevHitPoint == firstHitPoint
    assert np.array_equal(res["prevHitPoint"][absorbed], res["firstHitPoint"][absorbed])


def test_slab_snell_and_slots():
    """Refraction (normal_shader.cu:191-282): slot +R^3 is the ray inside the slab, slot +2R^3 the exit ray;
    entry obeys Snell with eta = 1/n; path rows are pre-filled as the reference does (:221-239)."""
    n_idx = 2.0
    targets, spec = scenes.slab(n=48, refr_index=n_idx)
    r = O.trace(targets, spec)
    res, rows = r["results"], r["targ_intersect"]
    R3, M, D = spec.rays, spec.slots, spec.depth_total
    assert (M, D) == (5, 4)
    s0, s1, s2 = res[:R3], res[R3:2 * R3], res[2 * R3:3 * R3]
    hit = s0["reflDepth"] >= 1
    assert hit.all()
    assert (s1["refrDepth"] == 1).all() and (s2["refrDepth"] == 2).all()
    assert (res[3 * R3:]["received"] == -1).all() and (res[3 * R3:]["rayLength"] == 0).all()   # slots never produced (Appendix B-Q5)
    assert (rows[R3:2 * R3] == 0).all()                                    # trapped ray: every column pre-filled
    assert (rows[2 * R3:3 * R3, :2] == 0).all()
    assert (rows[3 * R3:4 * R3, :3] == 0).all() and (rows[3 * R3:4 * R3, 3:] == -1).all()
    # Snell at entry: the interior segment runs from the first hit to the exit chain's first interior hit
    v, t = targets[0].verts, targets[0].tris
    tri = r["tri_path"][:R3, 0]
    p0, p1, p2 = v[t[tri, 0]], v[t[tri, 1]], v[t[tri, 2]]
    nrm = np.cross(p0 - p2, p1 - p0)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    d_in = s0["firstHitPoint"] / np.linalg.norm(s0["firstHitPoint"], axis=1)[:, None]
    # exit chain captured by receiver 0 without further hits: its prevHitPoint is the exit point
    ok = (s2["received"] == 0) & (s2["reflDepth"] == 0)
    assert ok.sum() > 100
    inside = s2["prevHitPoint"][ok] - s2["firstHitPoint"][ok]
    inside /= np.linalg.norm(inside, axis=1)[:, None]
    sin_i = np.linalg.norm(np.cross(d_in[ok], nrm[ok]), axis=1)
    sin_t = np.linalg.norm(np.cross(inside, nrm[ok]), axis=1)
    assert np.abs(sin_t * n_idx - sin_i).max() < 2e-5
    # transmitted power carries (1 - |coeff|) once per refraction while below the reflection budget (:245-246)
    assert (s2["power"][ok] > 0).all()
    bins, st = O.trace_bins(targets, spec, use_bvh=False)
    assert st["refracted"] == 2 * R3 and set(bins["rx"].tolist()) == {0, 1}


def test_total_internal_reflection_threshold():
    """optix::refract returns false when k = 1 - eta^2 (1 - c^2) < 0.  The plate's geometric normal
    (e1 x e0, triangle_mesh.cu:126) points along +x, i.e. WITH the incident ray, so refract() takes
    eta = ior = n2/n1 = 2 (Appendix B-Q18): rays steeper than asin(1/2) = 30 deg are not refracted."""
    plate = Target(np.array([[50.0, -400, -400], [50, 400, -400], [50, 400, 400], [50, -400, 400]]),
                   np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32), np.tile([[-1.0, 0, 0]], (4, 1)), 0.5, 2.0)
    spec = PulseSpec(grid=(1, 64, 1), max_refl=1, max_refr=2, tx_span=(2.4, 0.0, 0.0), rx=[], targ_vel=np.zeros((1, 3)))
    r = O.trace([plate], spec)
    R3 = spec.rays
    s0, s1 = r["results"][:R3], r["results"][R3:2 * R3]
    P = s0["firstHitPoint"]
    hit = s0["reflDepth"] == 1
    inc = np.degrees(np.arctan2(np.hypot(P[:, 1], P[:, 2]), P[:, 0]))
    refracted = s1["refrDepth"] == 1
    assert hit.all()
    assert refracted[inc < 29.9].all() and (~refracted[inc > 30.1]).all()


def test_direct_ray_power_and_window():
    """Direct rays (no target): power = 1/((4 pi)^2 R^2) (ray_tracer.cu:406), doppler = 0, path row all -1."""
    spec = PulseSpec(grid=(1, 32, 32), max_refl=2, max_refr=0, tx_span=(0.2, 0.2, 0.0),
                     rx=[O.rx_sphere_from_desc((200.0, 0, 0), math.pi, 0.0, 10.0, 2.0, 2.0)], targ_vel=np.zeros((0, 3)))
    r = O.trace([], spec)
    got = r["results"][r["results"]["received"] == 0]
    assert len(got) > 0 and (got["reflDepth"] == 0).all() and (got["doppler"] == 0).all()
    assert np.abs(got["power"] * FOUR_PI ** 2 * got["rayLength"] ** 2 - 1).max() < 1e-9
    assert (r["targ_intersect"] == -1).all()


def test_multi_capture_quirk():
    """Appendix B-Q8: no early exit in the receiver loop — a ray inside two windows is processed twice,
    `received` ends as the later index and the power is multiplied twice."""
    rx = [O.rx_sphere_from_desc((300.0, 0, 0), math.pi, 0.0, 20.0, 3.0, 3.0), O.rx_sphere_from_desc((500.0, 0, 0), math.pi, 0.0, 40.0, 3.0, 3.0)]
    plate = Target(np.array([[-50.0, -5, -5], [-50, 5, -5], [-50, 5, 5], [-50, -5, 5]]), np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32),
                   np.tile([[1.0, 0, 0]], (4, 1)), 1.0, 1.0)
    spec = PulseSpec(grid=(1, 8, 8), max_refl=1, max_refr=0, tx_origin=(0, 0, 0), tx_dir=(math.pi, 0.0), tx_span=(0.05, 0.05, 0.0), rx=rx,
                     targ_vel=np.zeros((1, 3)))
    r = O.trace([plate], spec)
    assert r["stats"]["multi_captured"] == spec.rays
    assert (r["results"]["received"] == 1).all()
