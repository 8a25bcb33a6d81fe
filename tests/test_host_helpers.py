"""Host-side helpers of the library (mesh generators, rigid rotation, receiver sphere) against the
oracle's restatement of /root/reference/ray_tracer.cpp:156-170, 226-504, 894-918 — bit exact."""
import math
import os

import numpy as np
import pytest

import oracle_api as O
from rts_b200 import lib


@pytest.mark.parametrize("args", [(8.0, 3.0, 3.0, 0.0, 0.0, 0.0), (2.5, 1.0, 7.0, 0.3, -0.2, 1.1)])
def test_rect_mesh(args):
    v, t, n = lib.rect_mesh(*args)
    ov, ot, on = O.rect_mesh(*args)
    assert v.shape == (8, 3) and t.shape == (12, 3) and n.shape == (12, 3)  # 12 FACE normals (ray_tracer.cpp:296)
    assert v.tobytes() == ov.tobytes() and np.array_equal(t, ot) and n.tobytes() == on.tobytes()
    # face normals are unit and orthogonal to their triangle
    e = v[t[:, 1]] - v[t[:, 0]]
    assert np.allclose(np.einsum("ij,ij->i", e, n), 0, atol=1e-12) and np.allclose(np.linalg.norm(n, axis=1), 1)


@pytest.mark.parametrize("subdivs,ntri", [(0, 20), (1, 80), (3, 1280)])
def test_sphere_mesh(subdivs, ntri):
    v, t, n = lib.sphere_mesh(subdivs, 2.0, 0.4, 0.1, -0.3)
    ov, ot, on = O.sphere_mesh(subdivs, 2.0, 0.4, 0.1, -0.3)
    assert len(t) == ntri and len(v) == 10 * 4 ** subdivs + 2
    assert v.tobytes() == ov.tobytes() and np.array_equal(t, ot) and n.tobytes() == on.tobytes()
    assert np.allclose(np.linalg.norm(v, axis=1), 2.0) and np.allclose(np.linalg.norm(n, axis=1), 1.0)


def test_file_mesh_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    tri = rng.normal(size=(5, 9))
    nrm = rng.normal(size=(5, 9))
    fmt = lambda row: "%.17g %.17g %.17g, %.17g %.17g %.17g, %.17g %.17g %.17g,\n" % tuple(row)
    vf, nf = tmp_path / "v.txt", tmp_path / "n.txt"
    vf.write_text("".join(fmt(r) for r in tri))
    nf.write_text("".join(fmt(r) for r in nrm))
    v, t, n = lib.file_mesh(vf, nf, 0.2, 0.0, 0.5)
    ov, ot, on = O.file_mesh(vf, nf, 0.2, 0.0, 0.5)
    assert len(v) == 15 and np.array_equal(t, np.arange(15, dtype=np.uint32).reshape(5, 3))  # 3 fresh vertices per triangle
    assert v.tobytes() == ov.tobytes() and np.array_equal(t, ot) and n.tobytes() == on.tobytes()
    v0, _, _ = lib.file_mesh(vf, nf)
    assert np.array_equal(v0, tri.reshape(15, 3))
    with pytest.raises(lib.RtsError):
        lib.file_mesh(tmp_path / "missing.txt", nf)


def test_rotation_matrix_float_angle_quirk():
    """vertex_rotation takes FLOAT angles: cos/sin are evaluated in float (ray_tracer.cpp:156-162)."""
    y, p, r = 0.3, -0.7, 1.9
    R = lib.rotation_matrix(y, p, r)
    cy, sy = np.float32(math.cos(np.float32(y))), None
    pts = np.eye(3)
    O.oracle().orc_vertex_rotation(pts.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)), 3,
                                   __import__("ctypes").c_float(y), __import__("ctypes").c_float(p), __import__("ctypes").c_float(r))
    assert np.array_equal(R.T, pts)                      # rows of rotated unit vectors = columns of R
    Rd = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6      # orthonormal to float precision only
    assert np.abs(R - R.astype(np.float32)).max() < 1e-6


def test_rx_sphere_from_desc():
    """ray_tracer.cpp:894-918: centre = position + r*(cos el cos az, cos el sin az, sin el) in float trig;
    the receiver position lies on the sphere and at the window centre."""
    for pos, az, el, r in [((0, 0, 0), 0.0, 0.0, 2.0), ((9000.0, 10.0, 1200.0), 3.0, -0.2, 250.0), ((-5, 3, 1), -1.1, 0.7, 20.0)]:
        a = lib.rx_sphere_from_desc(pos, az, el, r, 2.0, 1.0)
        b = O.rx_sphere_from_desc(pos, az, el, r, 2.0, 1.0)
        assert bytes(a) == bytes(b)
        c = np.array(a.centre[:])
        assert abs(np.linalg.norm(np.array(pos) - c) - r) < 1e-4 * r
        assert abs((a.max_theta - a.min_theta) - 2.0) < 1e-12 and abs((a.max_phi - a.min_phi) - 1.0) < 1e-12
    s = lib.rx_sphere_from_desc((0, 0, 0), 0.0, 0.0, 2.0, 2.0, 2.0)
    assert s.centre[:] == [2.0, 0.0, 0.0]                # Appendix C, C1
    assert abs(abs(0.5 * (s.min_theta + s.max_theta)) - math.pi) < 1e-6


@pytest.mark.parametrize("seed", range(12))
def test_random_helper_parameters_match_the_oracle(seed):
    """The library's host helpers and the oracle's restatement of the same reference lines agree bit for bit on random
    parameters (sizes from millimetres to kilometres, angles beyond one turn, negative radii)."""
    import ctypes as C
    rng = np.random.default_rng(600 + seed)
    w, h, d = (float(x) for x in 10.0 ** rng.uniform(-3, 3, 3))
    ypr = tuple(float(x) for x in rng.uniform(-8.0, 8.0, 3))
    a, b = lib.rect_mesh(w, h, d, *ypr), O.rect_mesh(w, h, d, *ypr)
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
    sub, rad = int(rng.integers(0, 4)), float(rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(-2, 2))
    a, b = lib.sphere_mesh(sub, rad, *ypr), O.sphere_mesh(sub, rad, *ypr)
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
    R = lib.rotation_matrix(*ypr)
    pts = np.eye(3)
    O.oracle().orc_vertex_rotation(pts.ctypes.data_as(C.POINTER(C.c_double)), 3, C.c_float(ypr[0]), C.c_float(ypr[1]), C.c_float(ypr[2]))
    assert np.array_equal(R.T, pts)
    pos = tuple(float(x) for x in rng.normal(0, 5000.0, 3))
    args = (pos, float(rng.uniform(-7, 7)), float(rng.uniform(-1.5, 1.5)), float(10.0 ** rng.uniform(-1, 3)), float(rng.uniform(0.01, 6.0)), float(rng.uniform(0.01, 3.0)))
    assert bytes(lib.rx_sphere_from_desc(*args)) == bytes(O.rx_sphere_from_desc(*args))


# ---- pinned to the reference's own code: oracle/_ref/libref_mesh.so is ray_tracer.cpp's helper functions compiled unmodified
needs_ref_mesh = pytest.mark.skipif(not os.path.exists(O.REF_MESH_LIB), reason="oracle/_ref/libref_mesh.so not built (make -C oracle ref)")


def _same(a, b):
    return all(x.shape == y.shape and x.tobytes() == y.tobytes() for x, y in zip(a, b))


@needs_ref_mesh
@pytest.mark.parametrize("seed", range(16))
def test_helpers_match_the_reference_functions(seed, tmp_path):
    """rts_rect_mesh / rts_sphere_mesh / rts_file_mesh / rts_rotation_matrix and the oracle's restatement, bit for bit
    against rect_mesh / sphere_mesh / file_mesh / vertex_rotation of /root/reference/ray_tracer.cpp itself."""
    rng = np.random.default_rng(900 + seed)
    w, h, d = (float(x) for x in 10.0 ** rng.uniform(-3, 3, 3))
    ypr = tuple(float(x) for x in rng.uniform(-8.0, 8.0, 3)) if seed else (0.0, 0.0, 0.0)
    want = O.ref_rect_mesh(w, h, d, *ypr)
    assert _same(lib.rect_mesh(w, h, d, *ypr), want) and _same(O.rect_mesh(w, h, d, *ypr), want)
    sub, rad = int(rng.integers(0, 4)), float(rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(-2, 2))
    want = O.ref_sphere_mesh(sub, rad, *ypr)
    assert _same(lib.sphere_mesh(sub, rad, *ypr), want) and _same(O.sphere_mesh(sub, rad, *ypr), want)
    want = O.ref_rotation_matrix(*ypr)
    assert np.array_equal(lib.rotation_matrix(*ypr), want) and np.array_equal(O.rotation_matrix(*ypr), want)
    n = int(rng.integers(1, 9))
    fmt = lambda row: "%.17g %.17g %.17g, %.17g %.17g %.17g, %.17g %.17g %.17g,\n" % tuple(row)
    vf, nf = tmp_path / "v.txt", tmp_path / "n.txt"
    vf.write_text("".join(fmt(r) for r in rng.normal(size=(n, 9)) * 10.0 ** rng.uniform(-2, 3)))
    nf.write_text("".join(fmt(r) for r in rng.normal(size=(n, 9))))
    want = O.ref_file_mesh(vf, nf, *ypr)
    assert _same(lib.file_mesh(vf, nf, *ypr), want) and _same(O.file_mesh(vf, nf, *ypr), want)
