"""Pins the oracle restatement to the reference's OWN source files.

oracle/_ref/libref_rts.so is /root/reference/{ray_tracer,triangle_mesh,normal_shader}.cu compiled
unmodified for the host against an OptiX emulation shim (oracle/ref_shim).  Live comparisons run when
that library is present (it is built in the dev container and travels to the GPU box); the committed
golden vectors under tests/golden/ were generated from it by tests/golden/make_golden.py and are
checked everywhere."""
import glob
import os

import numpy as np
import pytest

import oracle_api as O
import soups
from rts_b200 import scenes
from rts_b200.abi import RAY_RECORD

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "plate": lambda n: scenes.flat_plate(n=n, cubic=True),
    "plate_refl2": lambda n: scenes.flat_plate(n=n, cubic=True, max_refl=2),
    "trihedral": lambda n: scenes.trihedral(n=n, cubic=True),
    "trihedral_refl1": lambda n: scenes.trihedral(n=n, cubic=True, max_refl=1),
    "slab": lambda n: scenes.slab(n=n, cubic=True),
    "slab_interp_n13": lambda n: scenes.slab(n=n, cubic=True, interpolate=True, refr_index=1.3, max_refl=3),
    "slab_thin_refl0": lambda n: scenes.slab(n=n, cubic=True, thickness=0.004, max_refl=1),
    # random triangle soups (tests/soups.py): several targets, mixed materials and normal conventions, moving targets
    "soup_1": lambda n: soups.case(1, cubic_n=n),      # refraction, 4 targets
    "soup_4": lambda n: soups.case(4, cubic_n=n),      # smooth shading, reflections only
    "soup_18": lambda n: soups.case(18, cubic_n=n),    # refraction + smooth shading
}


def _records_equal(a, b):
    """Field-wise bit equality (the 144-byte record has alignment holes that serialisation may not keep)."""
    assert a.dtype.names == b.dtype.names and len(a) == len(b)
    for f in a.dtype.names:
        x, y = np.ascontiguousarray(a[f]), np.ascontiguousarray(b[f])
        assert x.tobytes() == y.tobytes(), f
    return True


def _same(a, b):
    assert _records_equal(a["results"], b["results"])
    assert np.array_equal(a["targ_intersect"], b["targ_intersect"])
    assert np.array_equal(a["tri_path"], b["tri_path"])
    assert a["rcs_angle"].tobytes() == b["rcs_angle"].tobytes()


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libref_rts.so not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_sources(name):
    targets, spec = CASES[name](12)
    a, b = O.trace(targets, spec), O.ref_trace(targets, spec)
    _same(a, b)
    assert a["stats"]["segments"] == b["segments"]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libref_rts.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(20))
def test_oracle_matches_reference_sources_on_random_soups(seed):
    """The restatement against the reference's own programs on inputs nobody arranged: random triangle soups."""
    targets, spec = soups.case(seed, cubic_n=8 + seed % 3)
    a, b = O.trace(targets, spec, use_bvh=False), O.ref_trace(targets, spec)
    _same(a, b)
    assert a["stats"]["segments"] == b["segments"]
    c = O.trace(targets, spec, use_bvh=True)            # and the oracle's BVH mode against both
    _same(c, b)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libref_rts.so not built (needs /root/reference)")
def test_bound_program_matches():
    """triangle_mesh.cu:204-233 (directed rounding) vs numpy nextafter."""
    t, _ = scenes.slab(n=4)
    ref = O.ref_bounds(t[0])
    v = t[0].verts[t[0].tris]                       # [T,3,3]
    lo, hi = v.min(axis=1), v.max(axis=1)
    lo32, hi32 = lo.astype(np.float32), hi.astype(np.float32)
    lo32 = np.where(lo32.astype(np.float64) > lo, np.nextafter(lo32, np.float32(-np.inf)), lo32)
    hi32 = np.where(hi32.astype(np.float64) < hi, np.nextafter(hi32, np.float32(np.inf)), hi32)
    assert np.array_equal(ref[:, :3], lo32) and np.array_equal(ref[:, 3:], hi32)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz"))))
def test_oracle_matches_golden_vectors(path):
    g = np.load(path)
    name, n = str(g["case"]), int(g["n"])
    targets, spec = CASES[name](n)
    a = O.trace(targets, spec)
    assert _records_equal(a["results"], g["results"])
    assert np.array_equal(a["targ_intersect"], g["targ_intersect"])
    assert np.array_equal(a["tri_path"], g["tri_path"])
    assert a["rcs_angle"].tobytes() == g["rcs_angle"].tobytes()


def test_golden_vectors_exist():
    assert len(glob.glob(os.path.join(GOLDEN, "ref_*.npz"))) >= 4
