"""Synthetic scenes for the BASELINE.json configs — benchmark and test inputs.

Every generator returns ``(targets, pulse_spec)`` built from plain numpy so that the very same
arrays can be handed to the CUDA library and to the CPU oracle.  Geometry that the reference
generates itself (``rect``/``sphere`` targets, rigid motion) comes from the library's host-side
generators (rts_b200.lib.rect_mesh / sphere_mesh / rotation_matrix), which restate
/root/reference/ray_tracer.cpp:156-170, 226-426.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .abi import PulseSpec, Target

C0 = 299792458.0


class _Helpers:
    """Where rect_mesh / sphere_mesh / rotation_matrix / rx_sphere_from_desc come from: the product library's host
    helpers by default (loaded on first use).  bench.py's reference arm installs the oracle's own restatements with
    set_helpers() so that the reference process never maps librts_b200.so."""
    _impl = None

    def __getattr__(self, name):
        if _Helpers._impl is None:
            from . import lib as _lib
            _Helpers._impl = _lib
        return getattr(_Helpers._impl, name)


lib = _Helpers()


def set_helpers(module) -> None:
    """Use `module` (rect_mesh, sphere_mesh, rotation_matrix, rx_sphere_from_desc) for the generated geometry."""
    _Helpers._impl = module


def _rx(position, azimuth, elevation, radius, theta_span=2.0, phi_span=2.0):
    return lib.rx_sphere_from_desc(position, azimuth, elevation, radius, theta_span, phi_span)


def flat_plate(n: int = 256, cubic: bool = False, max_refl: int = 1) -> Tuple[List[Target], PulseSpec]:
    """C1: 10 m x 10 m plate at x = 100 m, Tx and Rx at the origin (Appendix C)."""
    verts = np.array([[100, -5, -5], [100, 5, -5], [100, 5, 5], [100, -5, 5]], dtype=np.float64)
    tris = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    normals = np.tile(np.array([[-1.0, 0, 0]]), (4, 1))
    targets = [Target(verts, tris, normals, refl_coeff=1.0, refr_index=1.0)]
    grid = (n, n, n) if cubic else (1, n, n)
    spec = PulseSpec(grid=grid, max_refl=max_refl, max_refr=0, tx_origin=(0, 0, 0), tx_dir=(0.0, 0.0),
                     tx_span=(0.09, 0.09, 0.0), rx=[_rx((0, 0, 0), 0.0, 0.0, 2.0)])
    return targets, spec


def trihedral(n: int = 1000, cubic: bool = False, max_refl: int = 3) -> Tuple[List[Target], PulseSpec]:
    """C2: trihedral corner reflector, apex (60,60,60), three 10 m faces, 3-bounce retro-reflection."""
    a, s = 60.0, 10.0
    lo = a - s
    faces = []
    # plane x = a, y = a, z = a; each a quad split in two triangles; one target per face so that
    # the path rows distinguish the faces
    quads = [
        np.array([[a, lo, lo], [a, a, lo], [a, a, a], [a, lo, a]]),
        np.array([[lo, a, lo], [lo, a, a], [a, a, a], [a, a, lo]]),
        np.array([[lo, lo, a], [a, lo, a], [a, a, a], [lo, a, a]]),
    ]
    normals = [np.array([-1.0, 0, 0]), np.array([0, -1.0, 0]), np.array([0, 0, -1.0])]
    for q, nrm in zip(quads, normals):
        faces.append(Target(q, np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32), np.tile(nrm, (4, 1)), 1.0, 1.0))
    az, el = math.pi / 4, math.atan(1 / math.sqrt(2))
    grid = (n, n, n) if cubic else (1, n, n)
    spec = PulseSpec(grid=grid, max_refl=max_refl, max_refr=0, tx_origin=(0, 0, 0), tx_dir=(az, el),
                     tx_span=(0.08, 0.08, 0.0), rx=[_rx((0, 0, 0), az, el, 3.0)])
    return faces, spec


def slab(n: int = 64, cubic: bool = False, refr_index: float = 2.0, refl_coeff: float = 0.6, max_refl: int = 2,
         thickness: float = 1.0, tilt: float = 0.2, interpolate: bool = False) -> Tuple[List[Target], PulseSpec]:
    """Refraction KAT: a dielectric `rect` slab in front of the Tx, tilted so refraction bends the rays;
    one receiver behind the slab (exit rays), one at the Tx (reflections)."""
    v, t, fn = lib.rect_mesh(thickness, 30.0, 30.0, yaw=tilt, pitch=0.1, roll=0.0)
    v = v + np.array([50.0, 0.0, 0.0])
    targets = [Target(v, t, fn, refl_coeff=refl_coeff, refr_index=refr_index)]
    grid = (n, n, n) if cubic else (1, n, n)
    rx = [_rx((120.0, 0.0, 0.0), math.pi, 0.0, 30.0, 3.0, 3.0), _rx((-5.0, 0.0, 0.0), 0.0, 0.0, 20.0, 3.0, 3.0)]
    spec = PulseSpec(grid=grid, max_refl=max_refl, max_refr=2, interpolate_smooth=interpolate, tx_origin=(0, 0, 0),
                     tx_dir=(0.0, 0.0), tx_span=(0.2, 0.2, 0.0), rx=rx,
                     targ_vel=np.array([[3.0, -2.0, 1.0]]))
    return targets, spec


def direct_and_plate(n: int = 64, side: int = 1) -> Tuple[List[Target], PulseSpec]:
    """Direct-ray rule KAT (aggregation.cu:56, ray_tracer.cpp:1289-1294): one receiver that captures both direct rays
    and rays mirrored by a plate in the plane y = 30*side.  side=+1: the receiver's first received ray (lowest slot) is
    a direct one, so the direct bin emits its own response; side=-1: a reflected ray comes first and the direct rays'
    d_pathMatch collapses onto that path's representative."""
    y = 30.0 * side
    verts = np.array([[85.0, y, -6.0], [115.0, y, -6.0], [115.0, y, 6.0], [85.0, y, 6.0]])
    tris = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    normals = np.tile(np.array([[0.0, -float(side), 0.0]]), (4, 1))
    targets = [Target(verts, tris, normals, refl_coeff=0.8, refr_index=1.0)]
    spec = PulseSpec(grid=(1, n, n), max_refl=2, max_refr=0, tx_origin=(0, 0, 0), tx_dir=(0.0, 0.0), tx_span=(0.8, 0.2, 0.0),
                     rx=[_rx((200.0, 0.0, 0.0), math.pi, 0.0, 10.0, 2.0, 2.0)], targ_vel=np.array([[4.0, 1.0, 0.0]]))
    return targets, spec


def spheres(n: int = 96, interpolate: bool = True, max_refl: int = 2, max_refr: int = 2) -> Tuple[List[Target], PulseSpec]:
    """Two moving dielectric `sphere` targets (subdivided icosahedra, vertex normals = unit positions) in front of a
    metal back plate: exercises the barycentric normal interpolation of triangle_mesh.cu:186-190, Doppler from target
    velocities (normal_shader.cu:251-256, 302-314) and refraction through curved surfaces."""
    out = []
    for centre, radius, sub, refl, refr in (((40.0, -5.0, 1.0), 4.0, 3, 0.5, 1.5), ((55.0, 6.0, -2.0), 5.0, 2, 0.7, 2.2)):
        v, t, nrm = lib.sphere_mesh(sub, radius)
        out.append(Target(v + np.array(centre), t, nrm, refl_coeff=refl, refr_index=refr))
    v, t, fn = lib.rect_mesh(1.0, 60.0, 40.0)
    out.append(Target(v + np.array([90.0, 0.0, 0.0]), t, fn, refl_coeff=1.0, refr_index=1.0))
    rx = [_rx((0.0, 0.0, 0.0), 0.0, 0.0, 30.0, 2.5, 2.5), _rx((20.0, 50.0, 0.0), -1.2, 0.0, 25.0, 3.0, 3.0)]
    spec = PulseSpec(grid=(1, n, n), max_refl=max_refl, max_refr=max_refr, interpolate_smooth=interpolate, tx_origin=(0, 0, 0),
                     tx_dir=(0.0, 0.0), tx_span=(0.45, 0.3, 0.0), rx=rx,
                     targ_vel=np.array([[30.0, -12.0, 4.0], [-25.0, 8.0, -3.0], [0.0, 0.0, 0.0]]))
    return out, spec


# ---- C3: dielectric "ship" -------------------------------------------------------------------

def _superellipsoid(nu: int, nv: int, a: float, b: float, c: float, e1: float = 0.6, e2: float = 0.8):
    """File-style mesh (3 fresh vertices per triangle, per-vertex normals) of a superellipsoid."""
    u = np.linspace(-math.pi, math.pi, nu + 1)
    v = np.linspace(-math.pi / 2, math.pi / 2, nv + 1)
    U, V = np.meshgrid(u, v, indexing="ij")

    def spow(x, p):
        return np.sign(x) * np.abs(x) ** p

    X = a * spow(np.cos(V), e1) * spow(np.cos(U), e2)
    Y = b * spow(np.cos(V), e1) * spow(np.sin(U), e2)
    Z = c * spow(np.sin(V), e1)
    P = np.stack([X, Y, Z], axis=-1)
    # outward normal of a superellipsoid: gradient of the implicit function
    NX = spow(np.cos(V), 2 - e1) * spow(np.cos(U), 2 - e2) / a
    NY = spow(np.cos(V), 2 - e1) * spow(np.sin(U), 2 - e2) / b
    NZ = spow(np.sin(V), 2 - e1) / c
    N = np.stack([NX, NY, NZ], axis=-1)
    nrm = np.linalg.norm(N, axis=-1, keepdims=True)
    N = np.where(nrm > 0, N / np.maximum(nrm, 1e-300), np.array([0.0, 0.0, 1.0]))
    p00, p10, p11, p01 = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    n00, n10, n11, n01 = N[:-1, :-1], N[1:, :-1], N[1:, 1:], N[:-1, 1:]
    tv = np.concatenate([np.stack([p00, p10, p11], axis=2).reshape(-1, 3, 3), np.stack([p00, p11, p01], axis=2).reshape(-1, 3, 3)])
    tn = np.concatenate([np.stack([n00, n10, n11], axis=2).reshape(-1, 3, 3), np.stack([n00, n11, n01], axis=2).reshape(-1, 3, 3)])
    # drop degenerate triangles at the poles
    e0 = tv[:, 1] - tv[:, 0]
    e1_ = tv[:, 2] - tv[:, 0]
    area = np.linalg.norm(np.cross(e0, e1_), axis=1)
    keep = area > 1e-9
    tv, tn = tv[keep], tn[keep]
    verts = tv.reshape(-1, 3)
    normals = tn.reshape(-1, 3)
    tris = np.arange(len(verts), dtype=np.uint32).reshape(-1, 3)
    return verts, tris, normals


def ship(n: int = 4096, hull_res: int = 200, cubic: bool = False) -> Tuple[List[Target], PulseSpec]:
    """C3: ~100k-triangle dielectric ship (superellipsoid hull + box superstructure) on a sea plane,
    refraction on (max_refr=2), 4 receivers on an arc."""
    hv, ht, hn = _superellipsoid(hull_res + hull_res // 4, hull_res, 60.0, 8.0, 6.0)
    hull = Target(hv + np.array([0.0, 0.0, 3.0]), ht, hn, refl_coeff=0.6, refr_index=2.0)
    bv, bt, bn = lib.rect_mesh(30.0, 10.0, 8.0, yaw=0.05)
    box = Target(bv + np.array([-5.0, 0.0, 12.0]), bt, bn, refl_coeff=0.6, refr_index=2.0)
    sea_v = np.array([[-400.0, -400, 0], [400, -400, 0], [400, 400, 0], [-400, 400, 0]])
    sea = Target(sea_v, np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32), np.tile(np.array([[0.0, 0, 1]]), (4, 1)), 1.0, 1.0)
    targets = [hull, box, sea]
    tx = np.array([-2000.0, 0.0, 200.0])
    aim = np.array([0.0, 0.0, 5.0]) - tx
    az = math.atan2(aim[1], aim[0])
    el = math.atan2(aim[2], math.hypot(aim[0], aim[1]))
    rx = []
    for k in range(4):
        ang = math.radians(-15.0 + 10.0 * k)
        pos = np.array([-2000.0 * math.cos(ang), 2000.0 * math.sin(ang) + 40.0 * (k - 1.5), 200.0 + 60.0 * k])
        back = np.array([0.0, 0.0, 5.0]) - pos
        rx.append(_rx(pos, math.atan2(back[1], back[0]) + math.pi, -math.atan2(back[2], math.hypot(back[0], back[1])), 150.0,
                      2.4, 2.4))
    grid = (n, n, n) if cubic else (1, n, n)
    spec = PulseSpec(grid=grid, max_refl=3, max_refr=2, interpolate_smooth=True, tx_origin=tuple(tx), tx_dir=(az, el),
                     tx_span=(0.07, 0.012, 0.0), rx=rx, targ_vel=np.array([[8.0, 1.0, 0.0], [8.0, 1.0, 0.0], [0.0, 0, 0]]))
    return targets, spec


# ---- C4 / C5: 1M-triangle terrain + movers ----------------------------------------------------

def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _lattice(ix: np.ndarray, iy: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = _splitmix64(ix.astype(np.uint64) * np.uint64(0x1F1F1F1F1F1F1F1F) ^ _splitmix64(iy.astype(np.uint64) + np.uint64(seed)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def value_noise(x: np.ndarray, y: np.ndarray, seed: int, octaves: int = 5, base: float = 512.0) -> np.ndarray:
    """5-octave value noise in [0,1), lattice values from splitmix64."""
    out = np.zeros_like(x, dtype=np.float64)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        cell = base / (2 ** o)
        fx, fy = x / cell, y / cell
        ix, iy = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
        tx, ty = fx - ix, fy - iy
        tx, ty = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
        off = 1000003 * (o + 1)
        v00 = _lattice(ix + off, iy + off, seed)
        v10 = _lattice(ix + 1 + off, iy + off, seed)
        v01 = _lattice(ix + off, iy + 1 + off, seed)
        v11 = _lattice(ix + 1 + off, iy + 1 + off, seed)
        out += amp * ((v00 * (1 - tx) + v10 * tx) * (1 - ty) + (v01 * (1 - tx) + v11 * tx) * ty)
        tot += amp
        amp *= 0.5
    return out / tot


def terrain_mesh(cells_x: int = 1000, cells_y: int = 500, cell: float = 4.0, amplitude: float = 30.0,
                 seed: int = 0x52545301) -> Target:
    xs = np.arange(cells_x + 1) * cell
    ys = (np.arange(cells_y + 1) - cells_y / 2) * cell
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    Z = amplitude * value_noise(X, Y, seed)
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    idx = np.arange((cells_x + 1) * (cells_y + 1), dtype=np.uint32).reshape(cells_x + 1, cells_y + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    tris = np.concatenate([np.stack([a, b, c], axis=-1).reshape(-1, 3), np.stack([a, c, d], axis=-1).reshape(-1, 3)])
    # vertex normals from central differences
    gx = np.gradient(Z, cell, axis=0)
    gy = np.gradient(Z, cell, axis=1)
    N = np.stack([-gx, -gy, np.ones_like(Z)], axis=-1)
    N /= np.linalg.norm(N, axis=-1, keepdims=True)
    return Target(verts, tris.astype(np.uint32), N.reshape(-1, 3), refl_coeff=0.3, refr_index=1.0)


@dataclass
class MovingScene:
    """A scene whose targets move per pulse: base meshes + per-pulse poses (ray_tracer.cpp:936-1014)."""
    base: List[Target]
    spec: PulseSpec
    positions0: np.ndarray                 # [K,3] position at t = 0
    velocities: np.ndarray                 # [K,3] m/s
    rot_rates: np.ndarray                  # [K,3] yaw/pitch/roll rates rad/s (0 = not rotating)
    pri: float = 1e-3
    sample_time: float = 1e-3              # 1 / cw_sample_rate (ray_tracer.cpp:647)

    def poses(self, pulse: int):
        """(rotations, translations) of pulse k: rotation only for rotating targets at t > start."""
        t = pulse * self.pri
        rots: List[Optional[np.ndarray]] = []
        for k in range(len(self.base)):
            if np.any(self.rot_rates[k] != 0) and t > 0:
                y, p, r = (self.rot_rates[k] * t).tolist()
                rots.append(lib.rotation_matrix(y, p, r))
            else:
                rots.append(None)
        trans = self.positions0 + self.velocities * t
        return rots, trans

    def targ_vel(self, pulse: int) -> np.ndarray:
        """(p(t+dt) - p(t)) / dt, ray_tracer.cpp:1144-1145."""
        t = pulse * self.pri
        p0 = self.positions0 + self.velocities * t
        p1 = self.positions0 + self.velocities * (t + self.sample_time)
        return (p1 - p0) / self.sample_time

    def world_targets(self, pulse: int) -> List[Target]:
        """World-space meshes of pulse k computed on the host with the reference's operation order
        (what the oracle consumes): ((0 + R0*v0) + R1*v1) + R2*v2, then + t."""
        rots, trans = self.poses(pulse)
        out = []
        for k, b in enumerate(self.base):
            v, nrm = b.verts, b.normals
            if rots[k] is not None:
                R = rots[k]
                v = np.stack([((0.0 + R[i, 0] * v[:, 0]) + R[i, 1] * v[:, 1]) + R[i, 2] * v[:, 2] for i in range(3)], axis=1)
                nrm = np.stack([((0.0 + R[i, 0] * nrm[:, 0]) + R[i, 1] * nrm[:, 1]) + R[i, 2] * nrm[:, 2] for i in range(3)], axis=1)
            v = v + trans[k][None, :]
            out.append(Target(v, b.tris, nrm, b.refl_coeff, b.refr_index))
        return out

    def spec_for(self, pulse: int) -> PulseSpec:
        s = PulseSpec(**{**self.spec.__dict__})
        s.targ_vel = self.targ_vel(pulse)
        return s


def terrain_scene(n: int = 4096, cells_x: int = 1000, cells_y: int = 500, n_rx: int = 1, movers: int = 16,
                  nz: Optional[int] = None, seed: int = 0x52545301, extent_x: float = 4000.0) -> MovingScene:
    """C4/C5: terrain patch of extent_x by extent_x*cells_y/cells_x metres (static target 0; 1000 x 500 cells of
    4 m = 1,000,000 triangles by default) + boxes and spheres moving over it."""
    cell = extent_x / cells_x
    terr = terrain_mesh(cells_x, cells_y, cell, 30.0, seed)
    base = [terr]
    K = movers + 1
    pos0 = np.zeros((K, 3))
    vel = np.zeros((K, 3))
    rates = np.zeros((K, 3))
    lx, ly = cells_x * cell, cells_y * cell
    rnd = _splitmix64(np.arange(1, 16 * K + 1, dtype=np.uint64) + np.uint64(seed + 1))
    u = (rnd >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    u = u.reshape(K, 16)
    for k in range(1, K):
        if k % 2 == 1:
            v, t, nrm = lib.rect_mesh(8.0, 3.0, 3.0, yaw=float(u[k, 10] * 3.0))
        else:
            v, t, nrm = lib.sphere_mesh(3, 2.0)
        base.append(Target(v, t, nrm, refl_coeff=0.9, refr_index=1.0))
        pos0[k] = [lx * (0.2 + 0.6 * u[k, 0]), ly * (-0.4 + 0.8 * u[k, 1]), 20.0 + 80.0 * u[k, 2] + 30.0]
        vel[k] = (u[k, 3:6] - 0.5) * 100.0
        if k % 4 == 0:
            rates[k] = (u[k, 6:9] - 0.5) * 20.0
    tx = np.array([-3000.0, 0.0, 800.0])
    aim = np.array([lx / 2, 0.0, 0.0]) - tx
    az = math.atan2(aim[1], aim[0])
    el = math.atan2(aim[2], math.hypot(aim[0], aim[1]))
    rx = []
    for j in range(n_rx):
        ang = (j - (n_rx - 1) / 2) * 0.12
        pos = np.array([lx + 5000.0 * math.cos(ang), 5000.0 * math.sin(ang), 1500.0 + (1500.0 * j / max(1, n_rx - 1) if n_rx > 1 else 0.0)])
        rx.append(_rx(pos, math.pi + ang, 0.0, 250.0, 2.0, 2.0))
    grid = (1, n, nz if nz is not None else n)
    spec = PulseSpec(grid=grid, max_refl=3, max_refr=0, interpolate_smooth=False, tx_origin=tuple(tx), tx_dir=(az, el),
                     tx_span=(0.36, 0.14, 0.0), rx=rx, targ_vel=np.zeros((K, 3)))
    return MovingScene(base=base, spec=spec, positions0=pos0, velocities=vel, rot_rates=rates)
