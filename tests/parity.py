"""Shared comparison helpers for the parity tests (CUDA library vs CPU oracle)."""
from __future__ import annotations

import numpy as np

import oracle_api as O
from rts_b200 import lib as L

#: relative tolerance of the per-bin fp64 sums (BASELINE.json acceptance: 1e-5)
SUM_RTOL = 1e-5
#: rcs angles come from fp64 atan2 (CUDA libdevice vs glibc: a few ulp)
RCS_ATOL = 1e-12


def compare_records(gpu, orc, spec, label=""):
    """gpu = (results, targ_intersect, rcs_angle, tri_path) from Engine.records(); orc = oracle_api.trace() dict.
    Returns a dict of mismatch counts; rays flagged as receiver-window edge cases by the oracle are
    excluded from the `received`-dependent fields and counted separately."""
    res, ti, rcs, tp = gpu
    R3, M = spec.rays, spec.slots
    edge = orc["edge"]
    window_edge = (edge & O.EDGE_WINDOW) != 0
    slot_edge = np.tile(window_edge, M)
    out = {"label": label, "rays": int(R3), "window_edge_rays": int(window_edge.sum()),
           "tri_edge_rays": int(((edge & O.EDGE_TRI) != 0).sum()), "tie_rays": int(((edge & O.EDGE_TIE) != 0).sum())}
    o = orc["results"]
    out["tri_path_mismatch"] = int((tp != orc["tri_path"]).any(axis=1).sum())
    out["targ_intersect_mismatch"] = int((ti != orc["targ_intersect"]).any(axis=1).sum()) if spec.depth_total else 0
    for f in ("reflDepth", "refrDepth"):
        out[f + "_mismatch"] = int((res[f] != o[f]).sum())
    # fields independent of capture: bit-exact everywhere
    for f in ("firstHitPoint", "prevHitPoint"):
        out[f + "_mismatch"] = int((res[f].view(np.uint64) != o[f].view(np.uint64)).any(axis=1).sum())
    ok = ~slot_edge
    out["received_mismatch"] = int((res["received"][ok] != o["received"][ok]).sum())
    same_rx = ok & (res["received"] == o["received"])
    for f in ("rayLength", "power", "doppler"):
        out[f + "_mismatch"] = int((res[f][same_rx].view(np.uint64) != o[f][same_rx].view(np.uint64)).sum())
    for f in ("maxRayIndex", "end"):
        out[f + "_mismatch"] = int((res[f] != o[f]).sum())
    out["untouched_mismatch"] = int((res["rayDirection"] != o["rayDirection"]).any(axis=1).sum() + (res["refrIndex"] != o["refrIndex"]).any(axis=1).sum())
    if rcs is not None and spec.depth_total:
        d = np.abs(rcs - orc["rcs_angle"])
        out["rcs_angle_max_abs_diff"] = float(d.max()) if d.size else 0.0
    return out


def assert_records_equal(cmp):
    bad = {k: v for k, v in cmp.items() if k.endswith("_mismatch") and v}
    assert not bad, f"{cmp['label']}: {bad} (edge rays: window={cmp['window_edge_rays']} tri={cmp['tri_edge_rays']})"
    if "rcs_angle_max_abs_diff" in cmp:
        assert cmp["rcs_angle_max_abs_diff"] <= RCS_ATOL, cmp


def compare_bins(gbins, obins, rtol=SUM_RTOL):
    """Both arrays sorted by (rx, path). Returns dict with key/count equality and max relative errors."""
    out = {"n_gpu": len(gbins), "n_oracle": len(obins)}
    if len(gbins) != len(obins):
        out["keys_equal"] = False
        return out
    keys_equal = bool(np.array_equal(gbins["rx"], obins["rx"]) and np.array_equal(gbins["path"], obins["path"]))
    out["keys_equal"] = keys_equal
    out["npath_equal"] = bool(np.array_equal(gbins["npath"], obins["npath"]))
    out["min_slot_equal"] = bool(np.array_equal(gbins["min_slot"], obins["min_slot"]))
    out["direct_equal"] = bool(np.array_equal(gbins["direct"], obins["direct"]))
    for f in ("sum_sqrt_power", "sum_delay", "sum_phase", "sum_doppler", "power", "delay", "phase", "doppler"):
        a, b = gbins[f], obins[f]
        denom = np.maximum(np.abs(b), 1e-300)
        rel = np.abs(a - b) / denom
        rel = np.where((a == 0) & (b == 0), 0.0, rel)
        rel = np.where(np.isnan(a) & np.isnan(b), 0.0, rel)    # sqrt of a negative power (negative reflection coefficient) on both sides
        out[f + "_max_rel"] = float(rel.max()) if len(rel) else 0.0
    return out


def assert_bins_close(cmp, rtol=SUM_RTOL, exact_counts=True):
    assert cmp.get("keys_equal"), cmp
    if exact_counts:
        assert cmp["npath_equal"] and cmp["min_slot_equal"] and cmp["direct_equal"], cmp
    for k, v in cmp.items():
        if k.endswith("_max_rel"):
            # Doppler sums of static scenes are exactly 0 on both sides; otherwise relative
            assert v <= rtol, (k, v, cmp)


def run_gpu_records(engine, targets, spec, flags=None):
    engine.set_targets(targets)
    f = (L.RTS_OUT_RECORDS | L.RTS_OUT_BINS) if flags is None else flags
    stats = engine.trace(spec, f)
    return engine.records(), engine.bins(), stats


# ---- shards of large launches: compact records, and bins with the oracle's window-edge rays left out on both sides ----

def compare_shard_records(gpu, orc, spec, label=""):
    """gpu = Engine.records_shard(); orc = oracle_api.trace_shard(): both compact (ray k of the shard has slot s at
    k + s*n_shard).  Same fields and exclusions as compare_records."""
    res, ti, rcs, tp, n = gpu
    assert n == orc["n_shard"], (n, orc["n_shard"])
    M = spec.slots
    edge = orc["edge"]
    window_edge = (edge & O.EDGE_WINDOW) != 0
    slot_edge = np.tile(window_edge, M)
    out = {"label": label, "rays": int(n), "window_edge_rays": int(window_edge.sum()),
           "tri_edge_rays": int(((edge & O.EDGE_TRI) != 0).sum()), "tie_rays": int(((edge & O.EDGE_TIE) != 0).sum())}
    o = orc["results"]
    out["tri_path_mismatch"] = int((tp != orc["tri_path"]).any(axis=1).sum())
    out["targ_intersect_mismatch"] = int((ti != orc["targ_intersect"]).any(axis=1).sum()) if spec.depth_total else 0
    for f in ("reflDepth", "refrDepth"):
        out[f + "_mismatch"] = int((res[f] != o[f]).sum())
    for f in ("firstHitPoint", "prevHitPoint"):
        out[f + "_mismatch"] = int((res[f].view(np.uint64) != o[f].view(np.uint64)).any(axis=1).sum())
    ok = ~slot_edge
    out["received_mismatch"] = int((res["received"][ok] != o["received"][ok]).sum())
    same_rx = ok & (res["received"] == o["received"])
    for f in ("rayLength", "power", "doppler"):
        out[f + "_mismatch"] = int((res[f][same_rx].view(np.uint64) != o[f][same_rx].view(np.uint64)).sum())
    if rcs is not None and spec.depth_total:
        d = np.abs(rcs - orc["rcs_angle"])
        out["rcs_angle_max_abs_diff"] = float(d.max()) if d.size else 0.0
    return out


def specs_excluding(spec, flagged_k):
    """The shard of `spec` cut into sub-shards that leave out its shard-local rays `flagged_k`: a list of specs whose
    ray sets are disjoint and together are the shard minus those rays."""
    from rts_b200.abi import PulseSpec
    n = O.shard_size(spec)
    b, s = min(spec.ray_begin, spec.rays), spec.ray_stride or 1
    cuts = [-1] + sorted(int(k) for k in flagged_k) + [n]
    out = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        k0, k1 = lo + 1, hi            # shard-local [k0, k1)
        if k1 <= k0:
            continue
        sub = PulseSpec(**{**spec.__dict__})
        sub.ray_begin = b + k0 * s
        sub.ray_count = (k1 - k0 - 1) * s + 1
        sub.ray_stride = s
        out.append(sub)
    return out


def merge_bins(parts):
    """Bins of disjoint ray sets -> bins of their union: the five sums add, the representative slots take the minimum
    (the exchange step of the multi-GPU path, on finalised host arrays), myKernel2's means recomputed."""
    from rts_b200.abi import BIN_DTYPE
    acc = {}
    for bins in parts:
        for b in bins:
            key = (int(b["rx"]), tuple(int(x) for x in b["path"]))
            if key not in acc:
                acc[key] = b.copy()
            else:
                a = acc[key]
                for f in ("npath", "sum_sqrt_power", "sum_delay", "sum_phase", "sum_doppler"):
                    a[f] += b[f]
                a["min_slot"] = min(a["min_slot"], b["min_slot"])
                a["own_min_slot"] = min(a["own_min_slot"], b["own_min_slot"])
    out = np.zeros(len(acc), dtype=BIN_DTYPE)
    for i, key in enumerate(sorted(acc)):
        a = acc[key]
        a["power"] = (a["sum_sqrt_power"] / a["npath"]) ** 2
        a["delay"], a["phase"], a["doppler"] = a["sum_delay"] / a["npath"], a["sum_phase"] / a["npath"], a["sum_doppler"] / a["npath"]
        out[i] = a
    return out


def bins_excluding(engine, world_targets, spec, flagged_k, gpu_flags=0, use_bvh=True):
    """(gpu_bins, oracle_bins, n_excluded) of the spec's shard without the oracle's window-edge rays: both sides trace
    the same sub-shards in bins mode and merge them, so the comparison never has to be skipped."""
    subs = specs_excluding(spec, flagged_k)
    g, o = [], []
    for sub in subs:
        engine.trace(sub, L.RTS_OUT_BINS | gpu_flags)
        g.append(engine.bins().copy())
        o.append(O.trace_bins(world_targets, sub, use_bvh=use_bvh)[0])
    return merge_bins(g), merge_bins(o), len(list(flagged_k))


def assert_bins_match(engine, world_targets, spec, orc, use_bvh=True, gpu_flags=0, max_flagged=64):
    """Bins of the whole launch against the oracle's with the oracle's window-edge rays left out of BOTH sides (sub-shards
    around them): nothing is skipped, and the number of rays left out is bounded.  Returns that number."""
    flagged = np.nonzero((np.asarray(orc["edge"]) & O.EDGE_WINDOW) != 0)[0]
    assert len(flagged) <= max_flagged, f"{len(flagged)} window-edge rays"
    gb, ob, n_out = bins_excluding(engine, world_targets, spec, flagged, gpu_flags=gpu_flags, use_bvh=use_bvh)
    assert_bins_close(compare_bins(gb, ob))
    assert len(gb) == len(ob)
    return n_out
