"""Generates tests/golden/ref_*.npz from the reference's own sources (via oracle/_ref/libref_rts.so,
built by `make -C oracle ref` in a container that has /root/reference).  Run from the repo root:
    python tests/golden/make_golden.py
The vectors are the reference-shaped outputs (dbuf_results, dbuf_targ_intersect, dbuf_rcs_angle) plus the
shim's side-channel record of the winning triangle per closest-hit query."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_api as O  # noqa: E402
from test_oracle_reference import CASES  # noqa: E402

SIZES = {"plate": 8, "plate_refl2": 6, "trihedral": 8, "trihedral_refl1": 6, "slab": 6, "slab_interp_n13": 6, "slab_thin_refl0": 5,
         "soup_1": 7, "soup_4": 8, "soup_18": 7}

if __name__ == "__main__":
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    for name, n in SIZES.items():
        targets, spec = CASES[name](n)
        r = O.ref_trace(targets, spec)
        out = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(out, case=name, n=n, results=r["results"], targ_intersect=r["targ_intersect"],
                            rcs_angle=r["rcs_angle"], tri_path=r["tri_path"], segments=r["segments"])
        print(out, os.path.getsize(out), "bytes;", int((r["results"]["received"] >= 0).sum()), "received slots")
