"""ctypes mirror of include/rts_types.h (the POD carried across the C-ABI).

Field order, sizes and meaning follow include/rts_types.h exactly; tests/test_abi.py checks the
sizes against the compiled library.  The per-ray record is layout-identical to the reference's
``struct PerRayData`` (/root/reference/ray_tracer.h:13-28): size 144, align 16.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

RTS_MAX_DEPTH = 8
SCENE_EPS = np.float32(0.005)
SCENE_EPS_R = np.float32(0.005)

#: numpy view of rts_ray_record / PerRayData (ray_tracer.h:13-28)
RAY_RECORD = np.dtype(
    {
        "names": ["rayLength", "refrIndex", "reflDepth", "refrDepth", "maxRayIndex", "rayDirection",
                  "firstHitPoint", "prevHitPoint", "power", "doppler", "received", "end"],
        "formats": ["<f8", ("<f8", 2), "<u4", "<u4", "<u4", ("<f8", 3), ("<f8", 3), ("<f8", 3), "<f8", "<f8", "<i4", "u1"],
        "offsets": [0, 16, 32, 36, 40, 48, 72, 96, 120, 128, 136, 140],
        "itemsize": 144,
    }
)


class RtsTargetMesh(C.Structure):
    _fields_ = [
        ("n_verts", C.c_uint32), ("n_tris", C.c_uint32), ("n_normals", C.c_uint32), ("_pad", C.c_uint32),
        ("verts", C.POINTER(C.c_double)), ("tris", C.POINTER(C.c_uint32)), ("normals", C.POINTER(C.c_double)),
        ("refl_coeff", C.c_double), ("refr_index", C.c_double),
    ]


class RtsRxSphere(C.Structure):
    _fields_ = [("centre", C.c_double * 3), ("radius", C.c_double), ("min_theta", C.c_double),
                ("max_theta", C.c_double), ("min_phi", C.c_double), ("max_phi", C.c_double)]


class RtsRxDesc(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("azimuth", C.c_double), ("elevation", C.c_double),
                ("radius", C.c_double), ("theta_span", C.c_double), ("phi_span", C.c_double)]


class RtsPulse(C.Structure):
    _fields_ = [
        ("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32),
        ("max_refl", C.c_uint32), ("max_refr", C.c_uint32), ("interpolate_smooth", C.c_int32),
        ("tx_origin", C.c_double * 3), ("tx_dir", C.c_double * 2), ("tx_span", C.c_double * 3),
        ("cspeed", C.c_double), ("carrier", C.c_double),
        ("n_rx", C.c_uint32), ("n_targets", C.c_uint32),
        ("rx", C.POINTER(RtsRxSphere)), ("targ_vel", C.POINTER(C.c_double)),
        ("ray_begin", C.c_uint64), ("ray_count", C.c_uint64), ("ray_stride", C.c_uint64),
        ("targ_rcs", C.POINTER(C.c_double)), ("gain_tx", C.c_double), ("gain_rx", C.c_double),
    ]


class RtsBin(C.Structure):
    _fields_ = [
        ("rx", C.c_int32), ("path", C.c_int32 * RTS_MAX_DEPTH), ("direct", C.c_int32),
        ("npath", C.c_double), ("sum_sqrt_power", C.c_double), ("sum_delay", C.c_double),
        ("sum_phase", C.c_double), ("sum_doppler", C.c_double), ("min_slot", C.c_uint64), ("own_min_slot", C.c_uint64),
        ("power", C.c_double), ("delay", C.c_double), ("phase", C.c_double), ("doppler", C.c_double),
    ]


BIN_DTYPE = np.dtype(
    [("rx", "<i4"), ("path", "<i4", (RTS_MAX_DEPTH,)), ("direct", "<i4"), ("npath", "<f8"),
     ("sum_sqrt_power", "<f8"), ("sum_delay", "<f8"), ("sum_phase", "<f8"), ("sum_doppler", "<f8"),
     ("min_slot", "<u8"), ("own_min_slot", "<u8"), ("power", "<f8"), ("delay", "<f8"), ("phase", "<f8"), ("doppler", "<f8")],
    align=True,
)
assert BIN_DTYPE.itemsize == C.sizeof(RtsBin), (BIN_DTYPE.itemsize, C.sizeof(RtsBin))


class RtsResponse(C.Structure):
    _fields_ = [("rx", C.c_int32), ("_pad", C.c_int32), ("slot", C.c_uint64), ("power", C.c_double), ("delay", C.c_double),
                ("doppler", C.c_double), ("phase", C.c_double)]


class RtsTable2d(C.Structure):
    """include/rts_types.h: rts_table2d."""
    _fields_ = [("n_az", C.c_uint32), ("n_el", C.c_uint32), ("az0", C.c_double), ("az_step", C.c_double), ("el0", C.c_double),
                ("el_step", C.c_double), ("values", C.POINTER(C.c_double))]


class RtsAntenna(C.Structure):
    """include/rts_types.h: rts_antenna."""
    _fields_ = [("gain", RtsTable2d), ("bore_az", C.c_double), ("bore_el", C.c_double), ("rate_az", C.c_double), ("rate_el", C.c_double),
                ("position", C.c_double * 3)]


@dataclass
class Table2d:
    """A callback of two angles sampled on a regular grid (values[i, j] = f(az0 + i az_step, el0 + j el_step))."""
    az0: float
    az_step: float
    el0: float
    el_step: float
    values: np.ndarray

    def c(self) -> "RtsTable2d":
        self.values = np.ascontiguousarray(self.values, dtype=np.float64)
        t = RtsTable2d()
        t.n_az, t.n_el = self.values.shape
        t.az0, t.az_step, t.el0, t.el_step = float(self.az0), float(self.az_step), float(self.el0), float(self.el_step)
        t.values = self.values.ctypes.data_as(C.POINTER(C.c_double))
        return t

    def __call__(self, az, el):
        """The library's bilinear interpolation, in numpy (the host callback of the two-phase path in the tests)."""
        n_az, n_el = self.values.shape
        u = np.clip((np.asarray(az, dtype=np.float64) - self.az0) / self.az_step, 0.0, n_az - 1.0)
        v = np.clip((np.asarray(el, dtype=np.float64) - self.el0) / self.el_step, 0.0, n_el - 1.0)
        i0 = np.minimum(u.astype(np.int64), n_az - 1); j0 = np.minimum(v.astype(np.int64), n_el - 1)
        i1 = np.minimum(i0 + 1, n_az - 1); j1 = np.minimum(j0 + 1, n_el - 1)
        f, g = u - i0, v - j0
        V = self.values
        return (V[i0, j0] * (1 - f) + V[i1, j0] * f) * (1 - g) + (V[i0, j1] * (1 - f) + V[i1, j1] * f) * g


@dataclass
class Antenna:
    """include/rts_types.h: rts_antenna (gain None: the scalar gain of the pulse applies)."""
    position: tuple
    bore_az: float = 0.0
    bore_el: float = 0.0
    rate_az: float = 0.0
    rate_el: float = 0.0
    gain: Optional[Table2d] = None

    def c(self) -> "RtsAntenna":
        a = RtsAntenna()
        if self.gain is not None:
            a.gain = self.gain.c()
        a.bore_az, a.bore_el, a.rate_az, a.rate_el = float(self.bore_az), float(self.bore_el), float(self.rate_az), float(self.rate_el)
        a.position = (C.c_double * 3)(*[float(x) for x in self.position])
        return a


RESPONSE_DTYPE = np.dtype([("rx", "<i4"), ("_pad", "<i4"), ("slot", "<u8"), ("power", "<f8"), ("delay", "<f8"), ("doppler", "<f8"), ("phase", "<f8")])


class RtsStats(C.Structure):
    _fields_ = [
        ("primary_rays", C.c_uint64), ("segments", C.c_uint64), ("hits", C.c_uint64), ("shaded_hits", C.c_uint64),
        ("captured", C.c_uint64), ("multi_captured", C.c_uint64), ("edge_rays", C.c_uint64), ("refracted", C.c_uint64),
        ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("waves", C.c_uint64), ("kept_reflections", C.c_uint64),
        ("n_bins", C.c_uint32), ("primary_projected", C.c_uint32),
        ("ms_update", C.c_float), ("ms_trace", C.c_float), ("ms_finalise", C.c_float), ("ms_total", C.c_float),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("_")}


# ----------------------------------------------------------------------------------------------
# numpy-side scene description, convertible to the C structs above


@dataclass
class Target:
    """One target's world-space mesh and material (rts_target_mesh)."""
    verts: np.ndarray      # [V,3] float64
    tris: np.ndarray       # [T,3] uint32
    normals: np.ndarray    # [Nn,3] float64 (per-vertex, or per-face when Nn > V)
    refl_coeff: float = 1.0
    refr_index: float = 1.0

    def __post_init__(self):
        self.verts = np.ascontiguousarray(self.verts, dtype=np.float64).reshape(-1, 3)
        self.tris = np.ascontiguousarray(self.tris, dtype=np.uint32).reshape(-1, 3)
        self.normals = np.ascontiguousarray(self.normals, dtype=np.float64).reshape(-1, 3)


@dataclass
class PulseSpec:
    """Everything one pulse launch needs (rts_pulse)."""
    grid: Sequence[int]
    max_refl: int
    max_refr: int = 0
    interpolate_smooth: bool = False
    tx_origin: Sequence[float] = (0.0, 0.0, 0.0)
    tx_dir: Sequence[float] = (0.0, 0.0)
    tx_span: Sequence[float] = (0.1, 0.1, 0.0)
    cspeed: float = 299792458.0
    carrier: float = 10e9
    rx: List[RtsRxSphere] = field(default_factory=list)
    targ_vel: Optional[np.ndarray] = None  # [K,3]
    ray_begin: int = 0
    ray_count: int = 0
    ray_stride: int = 0
    targ_rcs: Optional[np.ndarray] = None  # [K] scalar RCS per target (fused bins), None = 1
    gain_tx: float = 1.0
    gain_rx: float = 1.0

    @property
    def rays(self) -> int:
        return int(self.grid[0]) * int(self.grid[1]) * int(self.grid[2])

    @property
    def r_max(self) -> int:
        return 2 if self.max_refr > 0 else 0

    @property
    def depth_total(self) -> int:   # D, ray_tracer.cpp:655
        return self.max_refl + self.r_max

    @property
    def slots(self) -> int:         # M, ray_tracer.cpp:608-613
        return 1 + (self.max_refl + 1) + 1 if self.r_max == 2 else 1

    @property
    def ray_total(self) -> int:
        return self.rays * self.slots

    @property
    def tri_cols(self) -> int:      # W
        return self.max_refl + 3


class CScene:
    """Keeps the ctypes arrays of a list of Targets alive."""

    def __init__(self, targets: Sequence[Target]):
        self.targets = list(targets)
        self.n = len(self.targets)
        self.array = (RtsTargetMesh * max(1, self.n))()
        for i, t in enumerate(self.targets):
            m = self.array[i]
            m.n_verts, m.n_tris, m.n_normals = len(t.verts), len(t.tris), len(t.normals)
            m.verts = t.verts.ctypes.data_as(C.POINTER(C.c_double))
            m.tris = t.tris.ctypes.data_as(C.POINTER(C.c_uint32))
            m.normals = t.normals.ctypes.data_as(C.POINTER(C.c_double))
            m.refl_coeff, m.refr_index = float(t.refl_coeff), float(t.refr_index)

    @property
    def total_tris(self) -> int:
        return sum(len(t.tris) for t in self.targets)


class CPulse:
    """Keeps the ctypes arrays of a PulseSpec alive."""

    def __init__(self, spec: PulseSpec, n_targets: int):
        self.spec = spec
        p = RtsPulse()
        p.nx, p.ny, p.nz = (int(v) for v in spec.grid)
        p.max_refl, p.max_refr = int(spec.max_refl), int(spec.max_refr)
        p.interpolate_smooth = 1 if spec.interpolate_smooth else 0
        p.tx_origin = (C.c_double * 3)(*[float(v) for v in spec.tx_origin])
        p.tx_dir = (C.c_double * 2)(*[float(v) for v in spec.tx_dir])
        p.tx_span = (C.c_double * 3)(*[float(v) for v in spec.tx_span])
        p.cspeed, p.carrier = float(spec.cspeed), float(spec.carrier)
        self.rx = (RtsRxSphere * max(1, len(spec.rx)))(*spec.rx)
        p.n_rx = len(spec.rx)
        p.rx = C.cast(self.rx, C.POINTER(RtsRxSphere))
        vel = spec.targ_vel if spec.targ_vel is not None else np.zeros((n_targets, 3))
        self.vel = np.ascontiguousarray(vel, dtype=np.float64).reshape(-1, 3)
        assert self.vel.shape[0] == n_targets, "targ_vel must have one row per target"
        p.n_targets = n_targets
        p.targ_vel = self.vel.ctypes.data_as(C.POINTER(C.c_double))
        p.ray_begin, p.ray_count, p.ray_stride = int(spec.ray_begin), int(spec.ray_count), int(spec.ray_stride)
        self.rcs = None
        if spec.targ_rcs is not None:
            self.rcs = np.ascontiguousarray(spec.targ_rcs, dtype=np.float64).reshape(-1)
            assert self.rcs.shape[0] == n_targets, "targ_rcs must have one entry per target"
            p.targ_rcs = self.rcs.ctypes.data_as(C.POINTER(C.c_double))
        p.gain_tx, p.gain_rx = float(spec.gain_tx), float(spec.gain_rx)
        self.c = p
