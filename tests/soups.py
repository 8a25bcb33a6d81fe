"""Random triangle-soup scenes shared by the fuzz tests (GPU vs oracle, tests/test_gpu_fuzz.py) and by the pin of
the oracle to the reference's own sources (tests/test_oracle_reference.py, tests/golden/make_golden.py).  Everything is
a function of the seed."""
import math

import numpy as np

from rts_b200 import scenes
from rts_b200.abi import PulseSpec, Target


def soup(rng, n_tris, centre, spread, size, per_face):
    """n_tris random triangles around `centre`; vertex normals (one per vertex) or per-face normals (Nn = T > V is the
    reference's file-mesh convention, triangle_mesh.cu:180 — so per-face soups share vertices to keep V < T)."""
    if per_face:
        nv = max(3, n_tris // 2)
        verts = centre + rng.normal(0.0, spread, (nv, 3))
        tris = np.stack([rng.choice(nv, 3, replace=False) for _ in range(n_tris)]).astype(np.uint32)
        nrm = rng.normal(0.0, 1.0, (n_tris, 3))
    else:
        c = centre + rng.normal(0.0, spread, (n_tris, 1, 3))
        verts = (c + rng.normal(0.0, size, (n_tris, 3, 3))).reshape(-1, 3)
        tris = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
        nrm = rng.normal(0.0, 1.0, (3 * n_tris, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return verts, tris, nrm


def case(seed, cubic_n=None):
    rng = np.random.default_rng(1000 + seed)
    K = int(rng.integers(1, 5))
    refr = bool(rng.integers(0, 2))
    targets = []
    for k in range(K):
        centre = np.array([rng.uniform(60, 140), rng.uniform(-25, 25), rng.uniform(-25, 25)])
        v, t, n = soup(rng, int(rng.integers(8, 90)), centre, rng.uniform(4, 14), rng.uniform(2, 9), per_face=bool(rng.integers(0, 2)))
        refl = float(rng.choice([1.0, -1.0, 0.9, 0.6, -0.4, 0.0]))
        targets.append(Target(v, t, n, refl_coeff=refl, refr_index=float(rng.choice([1.0, 1.3, 2.0, 0.7]))))
    cubic = seed % 3 == 0
    n = int(rng.integers(9, 15)) if cubic else int(rng.integers(40, 72))
    if cubic_n is not None:                     # the reference launches a cubic grid (ray_tracer.cpp:1165): goldens and the shim
        cubic, n = True, int(cubic_n)
    tx = np.array([rng.uniform(-20, 10), rng.uniform(-15, 15), rng.uniform(-15, 15)])
    aim = np.array([100.0, 0.0, 0.0]) - tx
    az, el = math.atan2(aim[1], aim[0]), math.atan2(aim[2], math.hypot(aim[0], aim[1]))
    rx = [scenes._rx(tuple(rng.uniform(-30, 30, 3) + np.array([-10.0, 0, 0])), az + rng.uniform(-0.3, 0.3), el + rng.uniform(-0.3, 0.3),
                     float(rng.uniform(4, 25)), float(rng.uniform(0.5, 3.0)), float(rng.uniform(0.5, 3.0))) for _ in range(int(rng.integers(1, 4)))]
    spec = PulseSpec(grid=(n, n, n) if cubic else (1, n, n + int(rng.integers(0, 9))), max_refl=int(rng.integers(1, 4)) if (refr and seed % 2) else int(rng.integers(0, 4)), max_refr=2 if refr else 0,
                     interpolate_smooth=bool(rng.integers(0, 2)), tx_origin=tuple(tx), tx_dir=(az, el),
                     tx_span=(float(rng.uniform(0.3, 0.9)), float(rng.uniform(0.3, 0.9)), 0.0), rx=rx,
                     targ_vel=rng.normal(0.0, 40.0, (K, 3)) * (rng.random((K, 1)) < 0.6))
    return targets, spec
