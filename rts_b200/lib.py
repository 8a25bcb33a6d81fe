"""ctypes binding of librts_b200.so (include/rts_b200.h) — the product's only compute path.

There is deliberately no fallback: if the shared library is missing the import of this module
raises, and every compute call raises RtsError when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .abi import (Antenna, RtsAntenna, RtsTable2d, Table2d, BIN_DTYPE, RAY_RECORD, RESPONSE_DTYPE, CPulse, CScene, PulseSpec, RtsBin, RtsPulse, RtsResponse, RtsRxDesc,
                  RtsRxSphere, RtsStats, RtsTargetMesh, Target)

RTS_OUT_BINS = 1
RTS_OUT_RECORDS = 2
RTS_COUNT_NODES = 4
RTS_NO_FINALISE = 8
RTS_NO_RCS_ANGLES = 16
RTS_ASYNC = 32
RTS_NO_REUSE = 64
RTS_TABLES = 128

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTS_B200_LIB", os.path.join(_HERE, "librts_b200.so"))   # override: tuning builds only


class RtsError(RuntimeError):
    pass


class RtsPose(C.Structure):
    _fields_ = [("R", C.c_double * 9), ("t", C.c_double * 3), ("has_rotation", C.c_int32), ("_pad", C.c_int32)]


class RtsSizes(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("ray_total", C.c_uint64), ("depth_total", C.c_uint32), ("slots", C.c_uint32),
                ("tri_cols", C.c_uint32), ("_pad", C.c_uint32)]


class RtsBvhInfo(C.Structure):
    _fields_ = [("n_tris", C.c_uint32), ("n_nodes", C.c_uint32), ("root_is_leaf", C.c_uint32), ("max_leaf", C.c_uint32),
                ("scene_lo", C.c_float * 3), ("scene_hi", C.c_float * 3), ("ms_build", C.c_float), ("ms_refit", C.c_float),
                ("sah_cost", C.c_double), ("sah_at_build", C.c_double), ("builds", C.c_uint32), ("builder", C.c_uint32)]


#: every symbol include/rts_b200.h declares (tests/test_abi.py checks the library exports them all)
EXPORTS = [
    "rts_create", "rts_destroy", "rts_last_error", "rts_version", "rts_abi_sizes", "rts_set_stream", "rts_set_option",
    "rts_rx_sphere_from_desc", "rts_result_sizes", "rts_rect_mesh", "rts_sphere_mesh", "rts_file_mesh",
    "rts_rotation_matrix", "rts_scene_set_targets", "rts_scene_set_poses", "rts_scene_rebuild", "rts_scene_bvh_info",
    "rts_scene_get_world_vertices", "rts_scene_get_tri_bounds", "rts_scene_check_bvh", "rts_trace_pulse",
    "rts_sync", "rts_get_stats", "rts_get_wave_profile", "rts_get_split_profile", "rts_get_follow_profile", "rts_kernel_launches", "rts_probe_read_bandwidth", "rts_get_bins", "rts_get_bins_previous", "rts_get_responses", "rts_get_records", "rts_get_records_shard", "rts_get_received", "rts_bins_device", "rts_bins_compact_device", "rts_bins_load_compact", "rts_finalise_bins", "rts_aggregate", "rts_set_rcs_tables", "rts_set_antennas", "rts_comm_create", "rts_comm_ipc_handle", "rts_comm_local_ptr", "rts_comm_connect_ipc", "rts_comm_connect_ptrs", "rts_comm_allreduce_bins", "rts_comm_stats", "rts_comm_destroy",
]

_lib = None


def load() -> C.CDLL:
    """Load librts_b200.so (built by __graft_entry__.build() / make -C rts_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RtsError(f"{LIB_PATH} is missing: build it with `make -C rts_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, dbl, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_double, C.c_int32
    P = C.POINTER
    lib.rts_create.argtypes = [C.c_int, P(vp)]
    lib.rts_destroy.argtypes = [vp]
    lib.rts_destroy.restype = None
    lib.rts_last_error.restype = C.c_char_p
    lib.rts_version.restype = C.c_char_p
    lib.rts_abi_sizes.argtypes = [P(u32)]
    lib.rts_set_stream.argtypes = [vp, vp]
    lib.rts_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.rts_rx_sphere_from_desc.argtypes = [P(RtsRxDesc), P(RtsRxSphere)]
    lib.rts_rx_sphere_from_desc.restype = None
    lib.rts_result_sizes.argtypes = [P(RtsPulse), P(RtsSizes)]
    mesh_tail = [P(dbl), P(u32), P(u32), P(u32), P(dbl), P(u32)]
    lib.rts_rect_mesh.argtypes = [C.c_float] * 6 + mesh_tail
    lib.rts_sphere_mesh.argtypes = [u32] + [C.c_float] * 4 + mesh_tail
    lib.rts_file_mesh.argtypes = [C.c_char_p, C.c_char_p] + [C.c_float] * 3 + mesh_tail
    lib.rts_rotation_matrix.argtypes = [C.c_float] * 3 + [P(dbl)]
    lib.rts_rotation_matrix.restype = None
    lib.rts_scene_set_targets.argtypes = [vp, P(RtsTargetMesh), u32]
    lib.rts_scene_set_poses.argtypes = [vp, P(RtsPose), u32]
    lib.rts_scene_rebuild.argtypes = [vp]
    lib.rts_scene_bvh_info.argtypes = [vp, P(RtsBvhInfo)]
    lib.rts_scene_get_world_vertices.argtypes = [vp, u32, P(dbl)]
    lib.rts_scene_get_tri_bounds.argtypes = [vp, P(C.c_float)]
    lib.rts_scene_check_bvh.argtypes = [vp, P(u64)]
    lib.rts_trace_pulse.argtypes = [vp, P(RtsPulse), u32]
    lib.rts_sync.argtypes = [vp]
    lib.rts_get_stats.argtypes = [vp, P(RtsStats)]
    lib.rts_get_wave_profile.argtypes = [vp, u32, P(C.c_float), P(u64), P(u32)]
    lib.rts_get_split_profile.argtypes = [vp, P(C.c_float)]
    lib.rts_get_follow_profile.argtypes = [vp, P(C.c_float)]
    lib.rts_comm_create.argtypes = [vp, u32, u32, u64]
    lib.rts_set_rcs_tables.argtypes = [vp, P(RtsTable2d), u32]
    lib.rts_set_antennas.argtypes = [vp, P(RtsAntenna), P(RtsAntenna), u32]
    lib.rts_comm_ipc_handle.argtypes = [vp, vp]
    lib.rts_comm_local_ptr.argtypes = [vp, P(vp)]
    lib.rts_comm_connect_ipc.argtypes = [vp, vp]
    lib.rts_comm_connect_ptrs.argtypes = [vp, P(vp)]
    lib.rts_comm_allreduce_bins.argtypes = [vp]
    lib.rts_comm_destroy.argtypes = [vp]
    lib.rts_comm_stats.argtypes = [vp, P(dbl)]
    lib.rts_kernel_launches.argtypes = [vp, P(u64)]
    lib.rts_probe_read_bandwidth.argtypes = [vp, u64, u32, P(dbl)]
    lib.rts_get_bins.argtypes = [vp, P(RtsBin), u32, P(u32)]
    lib.rts_get_bins_previous.argtypes = [vp, P(RtsBin), u32, P(u32)]
    lib.rts_get_responses.argtypes = [vp, P(RtsResponse), u32, P(u32)]
    lib.rts_get_records.argtypes = [vp, vp, P(i32), P(dbl), P(i32)]
    lib.rts_get_records_shard.argtypes = [vp, P(u64), vp, P(i32), P(dbl), P(i32)]
    lib.rts_get_received.argtypes = [vp, u64, P(u64), P(u64), vp, P(i32), P(dbl)]
    lib.rts_bins_device.argtypes = [vp, P(vp), P(u64), P(vp), P(u64)]
    lib.rts_bins_compact_device.argtypes = [vp, P(vp), P(vp), P(vp), P(u32)]
    lib.rts_bins_load_compact.argtypes = [vp, vp, vp, vp, u32]
    lib.rts_finalise_bins.argtypes = [vp]
    lib.rts_aggregate.argtypes = [vp, vp, P(i32), u32, u32, dbl, dbl, P(dbl), P(dbl), P(dbl), P(dbl), P(dbl), P(i32)]
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise RtsError(f"rts error {rc}: {load().rts_last_error().decode()}")


# ---- pure-host helpers (no GPU needed) --------------------------------------------------------

def rx_sphere_from_desc(position, azimuth, elevation, radius, theta_span, phi_span) -> RtsRxSphere:
    """Receiver sphere centre and angular window, /root/reference/ray_tracer.cpp:894-918."""
    d = RtsRxDesc()
    d.position = (C.c_double * 3)(*[float(v) for v in position])
    d.azimuth, d.elevation, d.radius = float(azimuth), float(elevation), float(radius)
    d.theta_span, d.phi_span = float(theta_span), float(phi_span)
    out = RtsRxSphere()
    load().rts_rx_sphere_from_desc(C.byref(d), C.byref(out))
    return out


def _mesh_call(fn, *head):
    nv, nt, nn = C.c_uint32(), C.c_uint32(), C.c_uint32()
    null_d, null_u = C.POINTER(C.c_double)(), C.POINTER(C.c_uint32)()
    _check(fn(*head, null_d, C.byref(nv), null_u, C.byref(nt), null_d, C.byref(nn)))
    v = np.zeros((nv.value, 3)); t = np.zeros((nt.value, 3), dtype=np.uint32); n = np.zeros((nn.value, 3))
    _check(fn(*head, v.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nv), t.ctypes.data_as(C.POINTER(C.c_uint32)),
              C.byref(nt), n.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nn)))
    return v, t, n


def rect_mesh(w, h, d, yaw=0.0, pitch=0.0, roll=0.0):
    """ray_tracer.cpp:226-297 — returns (verts[8,3], tris[12,3], face_normals[12,3])."""
    return _mesh_call(load().rts_rect_mesh, C.c_float(w), C.c_float(h), C.c_float(d), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


def sphere_mesh(subdivs, radius, yaw=0.0, pitch=0.0, roll=0.0):
    """ray_tracer.cpp:300-426 — subdivided icosahedron, vertex normals = unit positions."""
    return _mesh_call(load().rts_sphere_mesh, C.c_uint32(subdivs), C.c_float(radius), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


def file_mesh(v_file, n_file, yaw=0.0, pitch=0.0, roll=0.0):
    """ray_tracer.cpp:429-504 — text mesh, one triangle per line."""
    return _mesh_call(load().rts_file_mesh, str(v_file).encode(), str(n_file).encode(), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


def rotation_matrix(yaw, pitch, roll) -> np.ndarray:
    """Rz*Ry*Rx with the reference's float angles, ray_tracer.cpp:156-162."""
    R = np.zeros(9)
    load().rts_rotation_matrix(C.c_float(yaw), C.c_float(pitch), C.c_float(roll), R.ctypes.data_as(C.POINTER(C.c_double)))
    return R.reshape(3, 3)


def result_sizes(spec: PulseSpec, n_targets: int = 0) -> RtsSizes:
    cp = CPulse(spec, n_targets)
    out = RtsSizes()
    _check(load().rts_result_sizes(C.byref(cp.c), C.byref(out)))
    return out


# ---- the engine -------------------------------------------------------------------------------

class Engine:
    """One GPU's tracer: scene + BVH + wavefront launch (opaque rts_engine handle)."""

    def __init__(self, device: int = 0):
        self._lib = load()
        self._h = C.c_void_p()
        _check(self._lib.rts_create(int(device), C.byref(self._h)))
        self._scene: Optional[CScene] = None
        self._pulse: Optional[CPulse] = None
        self._stream = None            # cudaStream_t the engine was last pointed at (None: its own)

    def close(self):
        if self._h:
            self._lib.rts_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # scene -------------------------------------------------------------------------------
    def set_targets(self, targets: Sequence[Target]):
        self._scene = CScene(targets)
        _check(self._lib.rts_scene_set_targets(self._h, self._scene.array, self._scene.n))

    @staticmethod
    def pack_poses(rotations: Sequence[Optional[np.ndarray]], translations: Sequence[Sequence[float]]):
        """The rts_pose array of a pulse as the C-ABI takes it (a host loop that knows its poses ahead prepares them once)."""
        n = len(translations)
        arr = (RtsPose * max(1, n))()
        for k in range(n):
            R = rotations[k]
            arr[k].has_rotation = 0 if R is None else 1
            Rm = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
            arr[k].R = (C.c_double * 9)(*Rm.reshape(-1))
            arr[k].t = (C.c_double * 3)(*[float(v) for v in translations[k]])
        return arr, n

    def set_poses(self, rotations: Sequence[Optional[np.ndarray]], translations: Sequence[Sequence[float]]):
        self.set_poses_packed(*self.pack_poses(rotations, translations))

    def set_poses_packed(self, arr, n: int):
        _check(self._lib.rts_scene_set_poses(self._h, arr, n))

    def rebuild(self):
        _check(self._lib.rts_scene_rebuild(self._h))

    def bvh_info(self) -> RtsBvhInfo:
        out = RtsBvhInfo()
        _check(self._lib.rts_scene_bvh_info(self._h, C.byref(out)))
        return out

    def world_vertices(self, target: int) -> np.ndarray:
        out = np.zeros((len(self._scene.targets[target].verts), 3))
        _check(self._lib.rts_scene_get_world_vertices(self._h, target, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def tri_bounds(self) -> np.ndarray:
        out = np.zeros((self._scene.total_tris, 6), dtype=np.float32)
        _check(self._lib.rts_scene_get_tri_bounds(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def check_bvh(self) -> int:
        v = C.c_uint64()
        _check(self._lib.rts_scene_check_bvh(self._h, C.byref(v)))
        return int(v.value)

    # pulse -------------------------------------------------------------------------------
    def trace(self, spec: PulseSpec, flags: int = RTS_OUT_BINS) -> Optional[dict]:
        """One pulse.  With RTS_ASYNC the call returns as soon as the pulse is enqueued (None); stats() / bins() /
        sync() wait for it."""
        return self.trace_prepared(self.prepare(spec), flags)

    def prepare(self, spec: PulseSpec) -> "CPulse":
        """The rts_pulse struct (and the host arrays it points to) of a pulse, built once for trace_prepared."""
        return CPulse(spec, self._scene.n if self._scene else 0)

    def trace_prepared(self, pulse: "CPulse", flags: int = RTS_OUT_BINS) -> Optional[dict]:
        self._pulse = pulse
        _check(self._lib.rts_trace_pulse(self._h, C.byref(pulse.c), int(flags)))
        if flags & RTS_ASYNC:
            return None
        return self.stats()

    def sync(self):
        _check(self._lib.rts_sync(self._h))

    def stats(self) -> dict:
        s = RtsStats()
        _check(self._lib.rts_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def wave_profile(self):
        """[(ms, segments)] per bounce wave of the last pulse (CUDA events on the engine's stream)."""
        ms = (C.c_float * 32)(); seg = (C.c_uint64 * 32)(); n = C.c_uint32()
        _check(self._lib.rts_get_wave_profile(self._h, 32, ms, seg, C.byref(n)))
        return [(float(ms[i]), int(seg[i])) for i in range(n.value)]

    def split_profile(self):
        """(ms of k_traverse, ms of k_shade_wave) of the last pulse's second wave; (0, 0) when it ran fused."""
        ms = (C.c_float * 2)()
        _check(self._lib.rts_get_split_profile(self._h, ms))
        return float(ms[0]), float(ms[1])

    def follow_profile(self) -> float:
        """ms of k_primary_follow (primary shading + first reflections in place) in the last pulse; 0 when it did not run."""
        ms = C.c_float()
        _check(self._lib.rts_get_follow_profile(self._h, C.byref(ms)))
        return float(ms.value)

    def kernel_launches(self) -> int:
        v = C.c_uint64()
        _check(self._lib.rts_kernel_launches(self._h, C.byref(v)))
        return int(v.value)

    def probe_read_bandwidth(self, nbytes: int, reps: int = 50) -> float:
        """GB/s of 128-bit loads over a buffer of nbytes streamed reps times (L2-resident below ~100 MB)."""
        v = C.c_double()
        _check(self._lib.rts_probe_read_bandwidth(self._h, nbytes, reps, C.byref(v)))
        return float(v.value)

    def bins(self) -> np.ndarray:
        n = C.c_uint32()
        cap = 1024                                   # one call in the common case; *n reports the total
        out = np.zeros(cap, dtype=BIN_DTYPE)
        _check(self._lib.rts_get_bins(self._h, out.ctypes.data_as(C.POINTER(RtsBin)), cap, C.byref(n)))
        if n.value > cap:
            out = np.zeros(n.value, dtype=BIN_DTYPE)
            _check(self._lib.rts_get_bins(self._h, out.ctypes.data_as(C.POINTER(RtsBin)), n.value, C.byref(n)))
        return out[: n.value]

    def bins_previous(self) -> np.ndarray:
        """Bins of the pulse before the last one enqueued (pipelined host loops: rts_get_bins_previous)."""
        n = C.c_uint32()
        out = np.zeros(256, dtype=BIN_DTYPE)
        _check(self._lib.rts_get_bins_previous(self._h, out.ctypes.data_as(C.POINTER(RtsBin)), 256, C.byref(n)))
        return out[: n.value]

    def responses(self) -> np.ndarray:
        """One entry per response the reference would emit (ray_tracer.cpp:1289-1320), sorted by representative slot."""
        n = C.c_uint32()
        _check(self._lib.rts_get_responses(self._h, None, 0, C.byref(n)))
        out = np.zeros(max(1, n.value), dtype=RESPONSE_DTYPE)
        _check(self._lib.rts_get_responses(self._h, out.ctypes.data_as(C.POINTER(RtsResponse)), n.value, C.byref(n)))
        return out[: n.value]

    def records(self, rcs=True, tri_path=True):
        spec = self._pulse.spec
        n, D, W = spec.ray_total, spec.depth_total, spec.tri_cols
        res = np.zeros(n, dtype=RAY_RECORD)
        ti = np.zeros((n, max(D, 1)), dtype=np.int32)
        rc = np.zeros((n, max(D, 1), 2)) if rcs else None
        tp = np.zeros((n, W), dtype=np.int32) if tri_path else None
        _check(self._lib.rts_get_records(
            self._h, res.ctypes.data_as(C.c_void_p), ti.ctypes.data_as(C.POINTER(C.c_int32)),
            rc.ctypes.data_as(C.POINTER(C.c_double)) if rcs else None,
            tp.ctypes.data_as(C.POINTER(C.c_int32)) if tri_path else None))
        return res, ti[:, :D], (rc[:, :D] if rcs else None), tp

    def records_shard(self, rcs=True, tri_path=True):
        """Records of the last pulse's shard only, compact: ray k of the shard has slot s at index k + s*n_shard.
        Returns (results, targ_intersect, rcs_angle, tri_path, n_shard)."""
        spec = self._pulse.spec
        ns = C.c_uint64()
        _check(self._lib.rts_get_records_shard(self._h, C.byref(ns), None, None, None, None))
        n, D, W = int(ns.value) * spec.slots, spec.depth_total, spec.tri_cols
        res = np.zeros(n, dtype=RAY_RECORD)
        ti = np.zeros((n, max(D, 1)), dtype=np.int32)
        rc = np.zeros((n, max(D, 1), 2)) if rcs else None
        tp = np.zeros((n, W), dtype=np.int32) if tri_path else None
        _check(self._lib.rts_get_records_shard(
            self._h, C.byref(ns), res.ctypes.data_as(C.c_void_p), ti.ctypes.data_as(C.POINTER(C.c_int32)),
            rc.ctypes.data_as(C.POINTER(C.c_double)) if rcs else None,
            tp.ctypes.data_as(C.POINTER(C.c_int32)) if tri_path else None))
        return res, ti[:, :D], (rc[:, :D] if rcs else None), tp, int(ns.value)

    def received(self):
        """Received rays only, in slot order: (results, targ_intersect [R,D], rcs_angle [R,D,2], slots [R])
        — ray_tracer.cpp:1190-1221 before the RCS / gain callbacks."""
        D = self._pulse.spec.depth_total
        n = C.c_uint64()
        _check(self._lib.rts_get_received(self._h, 0, C.byref(n), None, None, None, None))
        R = int(n.value)
        res = np.zeros(max(1, R), dtype=RAY_RECORD)
        ti = np.zeros((max(1, R), max(D, 1)), dtype=np.int32)
        rc = np.zeros((max(1, R), max(D, 1), 2))
        sl = np.zeros(max(1, R), dtype=np.uint64)
        if R:
            _check(self._lib.rts_get_received(self._h, R, C.byref(n), sl.ctypes.data_as(C.POINTER(C.c_uint64)), res.ctypes.data_as(C.c_void_p),
                                              ti.ctypes.data_as(C.POINTER(C.c_int32)), rc.ctypes.data_as(C.POINTER(C.c_double))))
        if D == 0:
            return res[:R], ti[:R, :0], rc[:R, :0], sl[:R]
        return res[:R], ti[:R], rc[:R], sl[:R]

    def bins_device(self):
        """(sums_ptr, n_doubles, mins_ptr, n_u64) of the raw bin accumulators, for an external all-reduce."""
        sp, mp, ns, nm = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(self._lib.rts_bins_device(self._h, C.byref(sp), C.byref(ns), C.byref(mp), C.byref(nm)))
        return sp.value, int(ns.value), mp.value, int(nm.value)

    def bins_compact_device(self):
        """(keys_ptr, sums_ptr, mins_ptr, n) of this GPU's occupied sparse bins as compact device arrays."""
        kp, sp, mp, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32()
        _check(self._lib.rts_bins_compact_device(self._h, C.byref(kp), C.byref(sp), C.byref(mp), C.byref(n)))
        return kp.value, sp.value, mp.value, int(n.value)

    def bins_load_compact(self, keys_ptr: int, sums_ptr: int, mins_ptr: int, n: int):
        _check(self._lib.rts_bins_load_compact(self._h, C.c_void_p(keys_ptr), C.c_void_p(sums_ptr), C.c_void_p(mins_ptr), int(n)))

    def finalise_bins(self):
        _check(self._lib.rts_finalise_bins(self._h))

    # ---- tabulated callbacks (include/rts_b200.h: rts_set_rcs_tables / rts_set_antennas; used by pulses traced with RTS_TABLES)
    def set_rcs_tables(self, tables):
        """tables: one Table2d (or None: scalar RCS) per target; None / [] clears."""
        if not tables:
            _check(self._lib.rts_set_rcs_tables(self._h, None, 0))
            return
        arr = (RtsTable2d * len(tables))(*[t.c() if t is not None else RtsTable2d() for t in tables])
        _check(self._lib.rts_set_rcs_tables(self._h, arr, len(tables)))

    def set_antennas(self, tx, rx):
        """tx: Antenna or None; rx: list of Antenna (one per receiver) or None / [] to clear."""
        if not rx:
            txc = tx.c() if tx is not None else None      # a transmitter antenna alone is refused by the library
            _check(self._lib.rts_set_antennas(self._h, C.byref(txc) if txc is not None else None, None, 0))
            return
        arr = (RtsAntenna * len(rx))(*[a.c() for a in rx])
        txc = tx.c() if tx is not None else None
        _check(self._lib.rts_set_antennas(self._h, C.byref(txc) if txc is not None else None, arr, len(rx)))

    # ---- peer-memory bin exchange (include/rts_b200.h: rts_comm_*) ----
    def comm_create(self, rank: int, world: int, max_bins: int):
        _check(self._lib.rts_comm_create(self._h, int(rank), int(world), int(max_bins)))

    def comm_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self._lib.rts_comm_ipc_handle(self._h, buf))
        return buf.raw

    def comm_local_ptr(self) -> int:
        p = C.c_void_p()
        _check(self._lib.rts_comm_local_ptr(self._h, C.byref(p)))
        return int(p.value)

    def comm_connect_ipc(self, handles: bytes):
        buf = C.create_string_buffer(bytes(handles), len(handles))
        _check(self._lib.rts_comm_connect_ipc(self._h, buf))

    def comm_connect_ptrs(self, ptrs):
        arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
        _check(self._lib.rts_comm_connect_ptrs(self._h, arr))

    def comm_allreduce_bins(self):
        _check(self._lib.rts_comm_allreduce_bins(self._h))

    def comm_stats(self) -> dict:
        out = (C.c_double * 3)()
        _check(self._lib.rts_comm_stats(self._h, out))
        return {"exchanges": int(out[0]), "wait_us": out[1] * 1e-3, "kernel_us": out[2] * 1e-3}

    def comm_destroy(self):
        _check(self._lib.rts_comm_destroy(self._h))

    def set_option(self, name: str, value: int):
        """Tuning / test switch (include/rts_b200.h: rts_set_option); RTS_<NAME> in the environment sets the initial value."""
        _check(self._lib.rts_set_option(self._h, name.encode(), int(value)))

    def set_stream(self, cuda_stream: int):
        """Run the engine's work on this cudaStream_t (0 / None: back to the engine's own stream); waits for earlier work."""
        _check(self._lib.rts_set_stream(self._h, C.c_void_p(cuda_stream or 0)))
        self._stream = cuda_stream or None

    def aggregate(self, rx_results: np.ndarray, rx_intersects: np.ndarray, cspeed: float, carrier: float, ray_total: int):
        """rs::kernel_wrapper contract (aggregation.cu:103-184): returns dict of the written-back arrays."""
        R = len(rx_results)
        D = rx_intersects.shape[1] if rx_intersects.ndim == 2 else 0
        res = np.ascontiguousarray(rx_results.copy())
        rows = np.ascontiguousarray(rx_intersects, dtype=np.int32)
        acc = {k: np.zeros(R) for k in ("npath", "power", "doppler", "delay", "phase")}
        pm = np.full(R, ray_total + 1, dtype=np.int32)  # ray_tracer.cpp:1271
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        _check(self._lib.rts_aggregate(self._h, res.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.POINTER(C.c_int32)), R, D,
                                       float(cspeed), float(carrier), dp(acc["npath"]), dp(acc["power"]), dp(acc["doppler"]),
                                       dp(acc["delay"]), dp(acc["phase"]), pm.ctypes.data_as(C.POINTER(C.c_int32))))
        acc["results"] = res
        acc["path_match"] = pm
        return acc
