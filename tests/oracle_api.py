"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when present, to the reference
sources compiled against the OptiX emulation shim (oracle/_ref/libref_rts.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py, __graft_entry__.smoke()
and the cpu_baseline / --impl reference legs of bench.py.  Nothing under rts_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rts_b200.abi import BIN_DTYPE, RAY_RECORD, RESPONSE_DTYPE, RtsResponse, CPulse, CScene, PulseSpec, RtsBin, RtsPulse, RtsRxDesc, RtsRxSphere, RtsStats, RtsTargetMesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libref_rts.so")
REF_AGG_LIB = os.path.join(ORACLE_DIR, "_ref", "libref_aggregation.so")
REF_MESH_LIB = os.path.join(ORACLE_DIR, "_ref", "libref_mesh.so")   # the reference's own mesh helpers (oracle/Makefile: ref)

EDGE_TRI, EDGE_TIE, EDGE_WINDOW, EDGE_TMIN = 1, 2, 4, 8

_oracle = None
_ref = None


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", ORACLE_DIR, "_ref/libref_rts.so"], check=True, capture_output=True)


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_LIB):
            build_oracle()
        lib = C.CDLL(ORACLE_LIB)
        P, vp, u32, u64, dbl, i32 = C.POINTER, C.c_void_p, C.c_uint32, C.c_uint64, C.c_double, C.c_int32
        lib.orc_trace.argtypes = [P(RtsTargetMesh), u32, P(RtsPulse), C.c_int, vp, P(i32), P(dbl), P(i32), P(C.c_uint8), P(RtsStats)]
        lib.orc_trace_bins.argtypes = [P(RtsTargetMesh), u32, P(RtsPulse), C.c_int, P(RtsBin), u32, P(u32), P(RtsStats)]
        lib.orc_rx_sphere_from_desc.argtypes = [P(RtsRxDesc), P(RtsRxSphere)]
        lib.orc_rx_sphere_from_desc.restype = None
        lib.orc_postprocess.argtypes = [vp, P(i32), u64, u32, dbl, dbl, P(dbl), dbl, vp, P(i32), P(u64), u64, P(u64)]
        agg = [vp, P(i32), u32, u32, dbl, dbl, P(dbl), P(dbl), P(dbl), P(dbl), P(dbl), P(i32)]
        lib.orc_aggregate_literal.argtypes = agg
        lib.orc_aggregate_binned.argtypes = agg
        lib.orc_unique_paths.argtypes = [P(i32), u32, P(i32)]
        lib.orc_unique_paths.restype = u32
        lib.orc_responses.argtypes = [vp, P(dbl), P(dbl), P(i32), P(u64), u32, P(RtsResponse), u32]
        lib.orc_responses.restype = u32
        tail = [P(dbl), P(u32), P(u32), P(u32), P(dbl), P(u32)]
        lib.orc_rect_mesh.argtypes = [C.c_float] * 6 + tail
        lib.orc_sphere_mesh.argtypes = [u32] + [C.c_float] * 4 + tail
        lib.orc_file_mesh.argtypes = [C.c_char_p, C.c_char_p] + [C.c_float] * 3 + tail
        lib.orc_vertex_rotation.argtypes = [P(dbl), u32, C.c_float, C.c_float, C.c_float]
        lib.orc_vertex_rotation.restype = None
        lib.orc_trace_shard.argtypes = lib.orc_trace.argtypes
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        lib.orc_set_num_threads.restype = None
        lib.orc_version.restype = C.c_char_p
        _oracle = lib
    return _oracle


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_LIB)
        P, vp, u32, u64, dbl, i32 = C.POINTER, C.c_void_p, C.c_uint32, C.c_uint64, C.c_double, C.c_int32
        lib.ref_trace.argtypes = [P(RtsTargetMesh), u32, P(RtsPulse), vp, P(i32), P(dbl), P(i32), P(u64)]
        lib.ref_bounds.argtypes = [P(RtsTargetMesh), P(C.c_float)]
        lib.ref_version.restype = C.c_char_p
        _ref = lib
    return _ref


def _alloc(spec: PulseSpec):
    n, D, W = spec.ray_total, spec.depth_total, spec.tri_cols
    return (np.zeros(n, dtype=RAY_RECORD), np.zeros((n, max(D, 1)), dtype=np.int32), np.zeros((n, max(D, 1), 2)),
            np.zeros((n, W), dtype=np.int32))


def trace(targets, spec: PulseSpec, use_bvh=False):
    """Oracle launch with reference-shaped outputs: dict(results, targ_intersect, rcs_angle, tri_path, edge, stats)."""
    cs, cp = CScene(targets), CPulse(spec, len(targets))
    res, ti, rcs, tp = _alloc(spec)
    edge = np.zeros(spec.rays, dtype=np.uint8)
    st = RtsStats()
    rc = oracle().orc_trace(cs.array, cs.n, C.byref(cp.c), int(use_bvh), res.ctypes.data_as(C.c_void_p),
                            ti.ctypes.data_as(C.POINTER(C.c_int32)), rcs.ctypes.data_as(C.POINTER(C.c_double)),
                            tp.ctypes.data_as(C.POINTER(C.c_int32)), edge.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st))
    assert rc == 0, rc
    D = spec.depth_total
    return dict(results=res, targ_intersect=ti[:, :D], rcs_angle=rcs[:, :D], tri_path=tp, edge=edge, stats=st.as_dict())


def shard_size(spec: PulseSpec) -> int:
    """Number of primary rays a spec's ray_begin / ray_count / ray_stride select (api.cu / shard_bounds)."""
    rays = spec.rays
    b = min(spec.ray_begin, rays)
    e = min(rays, spec.ray_begin + spec.ray_count) if spec.ray_count else rays
    s = spec.ray_stride or 1
    return max(0, (max(e, b) - b + s - 1) // s)


def trace_shard(targets, spec: PulseSpec, use_bvh=True, arrays=True):
    """Oracle launch of a shard of a large grid with compact outputs: ray k of the shard has slot s at k + s*n_shard.
    arrays=False: only the per-ray edge flags and the counters."""
    cs, cp = CScene(targets), CPulse(spec, len(targets))
    n_sh, M, D, W = shard_size(spec), spec.slots, spec.depth_total, spec.tri_cols
    if not arrays:
        edge = np.zeros(n_sh, dtype=np.uint8)
        st = RtsStats()
        rc = oracle().orc_trace_shard(cs.array, cs.n, C.byref(cp.c), int(use_bvh), None, None, None, None,
                                      edge.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st))
        assert rc == 0, rc
        return dict(edge=edge, stats=st.as_dict(), n_shard=n_sh)
    n = n_sh * M
    res = np.zeros(n, dtype=RAY_RECORD)
    ti = np.zeros((n, max(D, 1)), dtype=np.int32)
    rcs = np.zeros((n, max(D, 1), 2))
    tp = np.zeros((n, W), dtype=np.int32)
    edge = np.zeros(n_sh, dtype=np.uint8)
    st = RtsStats()
    rc = oracle().orc_trace_shard(cs.array, cs.n, C.byref(cp.c), int(use_bvh), res.ctypes.data_as(C.c_void_p),
                                  ti.ctypes.data_as(C.POINTER(C.c_int32)), rcs.ctypes.data_as(C.POINTER(C.c_double)),
                                  tp.ctypes.data_as(C.POINTER(C.c_int32)), edge.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st))
    assert rc == 0, rc
    return dict(results=res, targ_intersect=ti[:, :D], rcs_angle=rcs[:, :D], tri_path=tp, edge=edge, stats=st.as_dict(), n_shard=n_sh)


def trace_bins(targets, spec: PulseSpec, use_bvh=True, cap=1 << 16):
    cs, cp = CScene(targets), CPulse(spec, len(targets))
    bins = np.zeros(cap, dtype=BIN_DTYPE)
    n = C.c_uint32()
    st = RtsStats()
    rc = oracle().orc_trace_bins(cs.array, cs.n, C.byref(cp.c), int(use_bvh), bins.ctypes.data_as(C.POINTER(RtsBin)), cap,
                                 C.byref(n), C.byref(st))
    assert rc == 0 and n.value <= cap, (rc, n.value)
    return bins[: n.value], st.as_dict()


def ref_trace(targets, spec: PulseSpec):
    """The reference's own programs (cubic grid only) through the OptiX emulation shim."""
    assert spec.grid[0] == spec.grid[1] == spec.grid[2], "the reference launches a cubic grid"
    cs, cp = CScene(targets), CPulse(spec, len(targets))
    res, ti, rcs, tp = _alloc(spec)
    seg = C.c_uint64()
    rc = ref().ref_trace(cs.array, cs.n, C.byref(cp.c), res.ctypes.data_as(C.c_void_p), ti.ctypes.data_as(C.POINTER(C.c_int32)),
                         rcs.ctypes.data_as(C.POINTER(C.c_double)), tp.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(seg))
    assert rc == 0, rc
    D = spec.depth_total
    return dict(results=res, targ_intersect=ti[:, :D], rcs_angle=rcs[:, :D], tri_path=tp, segments=int(seg.value))


def ref_bounds(target):
    cs = CScene([target])
    out = np.zeros((len(target.tris), 6), dtype=np.float32)
    assert ref().ref_bounds(cs.array, out.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return out


def postprocess(results, targ_intersect, spec: PulseSpec, rcs_per_target=None, gain=1.0):
    n, D = len(results), spec.depth_total
    ti = np.ascontiguousarray(targ_intersect, dtype=np.int32)
    nrx = int((results["received"] >= 0).sum())
    rx_res = np.zeros(max(1, nrx), dtype=RAY_RECORD)
    rx_rows = np.zeros((max(1, nrx), max(D, 1)), dtype=np.int32)
    rx_slots = np.zeros(max(1, nrx), dtype=np.uint64)
    cnt = C.c_uint64()
    rcs_p = None
    if rcs_per_target is not None:
        rcs_arr = np.ascontiguousarray(rcs_per_target, dtype=np.float64)
        rcs_p = rcs_arr.ctypes.data_as(C.POINTER(C.c_double))
    rows_flat = np.zeros((max(1, nrx), D), dtype=np.int32) if D else np.zeros((max(1, nrx), 0), dtype=np.int32)
    rc = oracle().orc_postprocess(results.ctypes.data_as(C.c_void_p), ti.ctypes.data_as(C.POINTER(C.c_int32)) if D else None, n, D,
                                  spec.cspeed, spec.carrier, rcs_p, gain, rx_res.ctypes.data_as(C.c_void_p),
                                  rows_flat.ctypes.data_as(C.POINTER(C.c_int32)) if D else None,
                                  rx_slots.ctypes.data_as(C.POINTER(C.c_uint64)), nrx, C.byref(cnt))
    assert rc == 0 and cnt.value == nrx
    return rx_res[:nrx], rows_flat[:nrx], rx_slots[:nrx]


def aggregate(rx_results, rx_rows, spec: PulseSpec, literal=True, ray_total=None):
    R = len(rx_results)
    D = rx_rows.shape[1] if rx_rows.ndim == 2 else 0
    res = np.ascontiguousarray(rx_results.copy())
    rows = np.ascontiguousarray(rx_rows, dtype=np.int32)
    acc = {k: np.zeros(max(1, R)) for k in ("npath", "power", "doppler", "delay", "phase")}
    pm = np.full(max(1, R), (ray_total if ray_total is not None else spec.ray_total) + 1, dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    fn = oracle().orc_aggregate_literal if literal else oracle().orc_aggregate_binned
    rc = fn(res.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.POINTER(C.c_int32)), R, D, spec.cspeed, spec.carrier,
            dp(acc["npath"]), dp(acc["power"]), dp(acc["doppler"]), dp(acc["delay"]), dp(acc["phase"]),
            pm.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    out = {k: v[:R] for k, v in acc.items()}
    out["results"] = res
    out["path_match"] = pm[:R]
    return out


def unique_paths(path_match):
    pm = np.ascontiguousarray(path_match, dtype=np.int32)
    out = np.zeros(max(1, len(pm)), dtype=np.int32)
    n = oracle().orc_unique_paths(pm.ctypes.data_as(C.POINTER(C.c_int32)), len(pm), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out[:n]


def responses(agg, rx_slots):
    """ray_tracer.cpp:1289-1320 on the output of aggregate(): one response per unique path_match value."""
    res = np.ascontiguousarray(agg["results"])
    R = len(res)
    out = np.zeros(max(1, R), dtype=RESPONSE_DTYPE)
    slots = np.ascontiguousarray(rx_slots, dtype=np.uint64)
    dp = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_double))
    n = oracle().orc_responses(res.ctypes.data_as(C.c_void_p), dp(agg["delay"]), dp(agg["phase"]),
                               np.ascontiguousarray(agg["path_match"], dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int32)),
                               slots.ctypes.data_as(C.POINTER(C.c_uint64)), R, out.ctypes.data_as(C.POINTER(RtsResponse)), R)
    return out[:n]


def responses_from_bins(bins):
    """The same responses derived from (receiver, path) bins (rule documented at rts_get_responses)."""
    keep = [b for b in bins if not (b["direct"] and b["own_min_slot"] != b["min_slot"])]
    out = np.zeros(len(keep), dtype=RESPONSE_DTYPE)
    for i, b in enumerate(sorted(keep, key=lambda b: int(b["min_slot"]))):
        out[i]["rx"], out[i]["slot"] = b["rx"], b["min_slot"]
        for f in ("power", "delay", "doppler", "phase"):
            out[i][f] = b[f]
    return out


def _mesh(fn, *head):
    nv, nt, nn = C.c_uint32(), C.c_uint32(), C.c_uint32()
    nd, nu = C.POINTER(C.c_double)(), C.POINTER(C.c_uint32)()
    assert fn(*head, nd, C.byref(nv), nu, C.byref(nt), nd, C.byref(nn)) == 0
    v = np.zeros((nv.value, 3)); t = np.zeros((nt.value, 3), dtype=np.uint32); n = np.zeros((nn.value, 3))
    assert fn(*head, v.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nv), t.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(nt),
              n.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nn)) == 0
    return v, t, n


def rect_mesh(w, h, d, yaw=0.0, pitch=0.0, roll=0.0):
    return _mesh(oracle().orc_rect_mesh, C.c_float(w), C.c_float(h), C.c_float(d), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


def sphere_mesh(subdivs, radius, yaw=0.0, pitch=0.0, roll=0.0):
    return _mesh(oracle().orc_sphere_mesh, C.c_uint32(subdivs), C.c_float(radius), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


def file_mesh(v_file, n_file, yaw=0.0, pitch=0.0, roll=0.0):
    return _mesh(oracle().orc_file_mesh, str(v_file).encode(), str(n_file).encode(), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


_ref_mesh = None


def ref_mesh():
    """oracle/_ref/libref_mesh.so: ray_tracer.cpp's rect_mesh / sphere_mesh / file_mesh / vertex_rotation compiled unmodified."""
    global _ref_mesh
    if _ref_mesh is None:
        lib = C.CDLL(REF_MESH_LIB)
        P, u32, dbl = C.POINTER, C.c_uint32, C.c_double
        tail = [P(dbl), P(u32), P(u32), P(u32), P(dbl), P(u32)]
        lib.refm_rect_mesh.argtypes = [C.c_float] * 6 + tail
        lib.refm_sphere_mesh.argtypes = [u32] + [C.c_float] * 4 + tail
        lib.refm_file_mesh.argtypes = [C.c_char_p, C.c_char_p] + [C.c_float] * 3 + tail
        lib.refm_vertex_rotation.argtypes = [P(dbl), u32, C.c_float, C.c_float, C.c_float]
        lib.refm_vertex_rotation.restype = None
        _ref_mesh = lib
    return _ref_mesh


def ref_rect_mesh(w, h, d, yaw=0.0, pitch=0.0, roll=0.0):
    return _mesh(ref_mesh().refm_rect_mesh, C.c_float(w), C.c_float(h), C.c_float(d), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


def ref_sphere_mesh(subdivs, radius, yaw=0.0, pitch=0.0, roll=0.0):
    return _mesh(ref_mesh().refm_sphere_mesh, C.c_uint32(subdivs), C.c_float(radius), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


def ref_file_mesh(v_file, n_file, yaw=0.0, pitch=0.0, roll=0.0):
    return _mesh(ref_mesh().refm_file_mesh, str(v_file).encode(), str(n_file).encode(), C.c_float(yaw), C.c_float(pitch), C.c_float(roll))


def ref_rotation_matrix(yaw, pitch, roll) -> np.ndarray:
    e = np.eye(3)
    ref_mesh().refm_vertex_rotation(e.ctypes.data_as(C.POINTER(C.c_double)), 3, C.c_float(yaw), C.c_float(pitch), C.c_float(roll))
    return np.ascontiguousarray(e.T)


def rotation_matrix(yaw, pitch, roll) -> np.ndarray:
    """Rz*Ry*Rx with the reference's float angles (ray_tracer.cpp:156-162): the rotated unit vectors are R's columns."""
    e = np.eye(3)
    oracle().orc_vertex_rotation(e.ctypes.data_as(C.POINTER(C.c_double)), 3, C.c_float(yaw), C.c_float(pitch), C.c_float(roll))
    return np.ascontiguousarray(e.T)


def set_num_threads(n: int) -> int:
    """OpenMP threads of the following oracle runs; returns the count in effect."""
    oracle().orc_set_num_threads(int(n))
    return int(oracle().orc_num_threads())


def rx_sphere_from_desc(position, azimuth, elevation, radius, theta_span, phi_span) -> RtsRxSphere:
    d = RtsRxDesc()
    d.position = (C.c_double * 3)(*[float(v) for v in position])
    d.azimuth, d.elevation, d.radius = float(azimuth), float(elevation), float(radius)
    d.theta_span, d.phi_span = float(theta_span), float(phi_span)
    out = RtsRxSphere()
    oracle().orc_rx_sphere_from_desc(C.byref(d), C.byref(out))
    return out
