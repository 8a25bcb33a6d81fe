"""Sparse (hashed) receiver bins: the reference groups arbitrary path rows (aggregation.cu:43-57), so the fused bins must
not be limited to key spaces a dense table can hold.  Beyond 2^20 dense bins — or with option "hash_bins" — the bins live
in an open-addressing table keyed by rx * (K+1)^D + path key; emission, clearing and the multi-GPU exchange cost what is
occupied.  Checked here: equal to the dense table where both exist, equal to the oracle where only the sparse one can
run (many targets x deep paths x several receivers), table re-use across pulses, the two-rank exchange, table overflow."""
import math

import numpy as np
import pytest
import torch

import oracle_api as O
import parity
from rts_b200 import dist as rdist
from rts_b200 import lib as L
from rts_b200 import scenes
from rts_b200.abi import PulseSpec, Target

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng():
    with L.Engine(0) as e:
        yield e


@pytest.mark.parametrize("name", ["trihedral", "slab", "spheres", "terrain"])
def test_sparse_bins_equal_dense_bins(eng, name):
    if name == "trihedral":
        targets, spec = scenes.trihedral(n=300)
    elif name == "slab":
        targets, spec = scenes.slab(n=96)
    elif name == "spheres":
        targets, spec = scenes.spheres(n=96)
    else:
        ms = scenes.terrain_scene(n=256, cells_x=100, cells_y=50, movers=6, n_rx=3)
        targets, spec = ms.world_targets(2), ms.spec_for(2)
    eng.set_targets(targets)
    st_d = eng.trace(spec, L.RTS_OUT_BINS)
    dense, resp_d = eng.bins().copy(), eng.responses().copy()
    eng.set_option("hash_bins", 1)
    for rep in range(3):                      # the table is cleared by walking the previous pulse's occupied slots
        st_s = eng.trace(spec, L.RTS_OUT_BINS | (L.RTS_ASYNC if rep == 1 else 0))
        sparse = eng.bins()
        cmp = parity.compare_bins(sparse, dense)
        parity.assert_bins_close(cmp, rtol=1e-12)
        assert np.array_equal(sparse["own_min_slot"], dense["own_min_slot"])
    resp_s = eng.responses()
    assert len(resp_s) == len(resp_d) and np.array_equal(resp_s["slot"], resp_d["slot"])
    assert np.allclose(resp_s["power"], resp_d["power"], rtol=1e-12, equal_nan=True)
    st_s = eng.stats()
    for k in ("segments", "hits", "captured"):
        assert st_s[k] == st_d[k]
    eng.set_option("hash_bins", 0)
    eng.trace(spec, L.RTS_OUT_BINS)           # and back to the dense table over the same arrays
    parity.assert_bins_close(parity.compare_bins(eng.bins(), dense), rtol=1e-12)


def _many_targets(n_side=8, n=160):
    """64 small plates in a staggered grid above a dielectric slab, 6 receivers: 65^5 path keys x 6 receivers ~ 7e9 dense
    bins — only the sparse table can hold them."""
    targets = []
    base, _ = scenes.slab(n=8)
    targets.append(base[0])
    rng = np.random.default_rng(11)
    for i in range(n_side):
        for j in range(n_side):
            v, t, fn = L.rect_mesh(0.5, 4.0, 4.0, yaw=float(rng.uniform(-0.6, 0.6)), pitch=float(rng.uniform(-0.6, 0.6)))
            c = np.array([60.0 + 3.0 * ((i + j) % 3), -28.0 + 8.0 * i, -28.0 + 8.0 * j])
            targets.append(Target(v + c, t, fn, refl_coeff=0.8, refr_index=1.0))
    rx = [L.rx_sphere_from_desc((-40.0, 60.0 * math.cos(a), 60.0 * math.sin(a)), 0.0, 0.0, 30.0, 2.0, 2.0) for a in np.linspace(0, 5, 6)]
    spec = PulseSpec(grid=(1, n, n), max_refl=3, max_refr=2, tx_origin=(0.0, 0.0, 0.0), tx_dir=(0.0, 0.0), tx_span=(0.9, 0.9, 0.0),
                     rx=rx, targ_vel=np.zeros((len(targets), 3)))
    return targets, spec


def test_key_space_beyond_any_dense_table_matches_the_oracle(eng):
    targets, spec = _many_targets()
    assert (len(targets) + 1) ** spec.depth_total * len(spec.rx) > 1 << 30
    eng.set_targets(targets)
    st = eng.trace(spec, L.RTS_OUT_BINS | L.RTS_OUT_RECORDS)
    orc = O.trace(targets, spec, use_bvh=False)
    cmp = parity.compare_records(eng.records(), orc, spec, "many-targets")
    parity.assert_records_equal(cmp)
    flagged = np.nonzero((orc["edge"] & O.EDGE_WINDOW) != 0)[0]
    assert len(flagged) <= 8
    gb, ob, _ = parity.bins_excluding(eng, targets, spec, flagged, use_bvh=False)
    parity.assert_bins_close(parity.compare_bins(gb, ob))
    assert len(gb) > 10 and st["captured"] > 0


def test_two_rank_exchange_of_sparse_bins(eng):
    """The multi-GPU exchange on one GPU: each "rank" traces its round-robin share into the sparse table, hands out its
    occupied bins as compact arrays; the merged arrays (dist.merge_compact = the host statement of exchange_sparse) are
    loaded back and finalised: equal to the un-sharded launch."""
    ms = scenes.terrain_scene(n=192, cells_x=80, cells_y=40, movers=6, n_rx=3)
    eng.set_targets(ms.base)
    eng.set_poses(*ms.poses(2))
    spec = ms.spec_for(2)
    eng.set_option("hash_bins", 1)
    eng.trace(spec, L.RTS_OUT_BINS)
    full = eng.bins().copy()
    dev = torch.device("cuda:0")
    world, parts = 3, []
    for r in range(world):
        spec.ray_begin, spec.ray_count, spec.ray_stride = r, 0, world
        eng.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_FINALISE)
        parts.append(rdist.compact_bins_as_tensors(eng, dev))
        with pytest.raises(L.RtsError):
            eng.bins_device()                      # the dense accessor refuses a sparse pulse
    union, u_sums, u_mins = rdist.merge_compact(parts)
    torch.cuda.synchronize()
    eng.bins_load_compact(union.data_ptr(), u_sums.contiguous().data_ptr(), u_mins.data_ptr(), union.numel())
    eng.finalise_bins()
    merged = eng.bins()
    parity.assert_bins_close(parity.compare_bins(merged, full), rtol=1e-12)
    assert sum(p[0].numel() for p in parts) >= union.numel() == len(full)


def test_sparse_table_overflow_is_reported(eng):
    targets, spec = scenes.trihedral(n=200)
    eng.set_targets(targets)
    eng.set_option("hash_bins", 1)
    eng.set_option("hash_log2", 4)             # 16 slots
    eng.trace(spec, L.RTS_OUT_BINS)            # a handful of bins: fits
    few = len(eng.bins())
    assert 0 < few <= 16
    targets, spec = _many_targets()            # 23 distinct (receiver, path) bins
    eng.set_targets(targets)
    with pytest.raises(L.RtsError, match="dropped|overflow|capacity"):
        eng.trace(spec, L.RTS_OUT_BINS)
    eng.set_option("hash_log2", 22)
    eng.trace(spec, L.RTS_OUT_BINS)
    assert len(eng.bins()) > 16
