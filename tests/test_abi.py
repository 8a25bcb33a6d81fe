"""The C-ABI library loads without a GPU, exports every symbol include/rts_b200.h declares, its POD
sizes match the ctypes mirror, and compute entry points fail loudly when no device is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rts_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rts_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rts_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    l = lib.load()
    declared = _declared_symbols()
    assert set(declared) == set(lib.EXPORTS), (set(declared) ^ set(lib.EXPORTS))
    for name in declared:
        assert hasattr(l, name), f"librts_b200.so does not export {name}"


def test_kernel_wrapper_cpp_symbol_exported():
    """rs::kernel_wrapper keeps the reference's mangled name (aggregation.cuh:18-23)."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "_ZN2rs14kernel_wrapperEP10PerRayDataPijjjjddPdS3_S3_S3_S3_S2_" in out


def test_pod_sizes_match_ctypes():
    sizes = (C.c_uint32 * 10)()
    assert lib.load().rts_abi_sizes(sizes) == 0
    expect = [144, C.sizeof(abi.RtsTargetMesh), C.sizeof(abi.RtsRxSphere), C.sizeof(abi.RtsRxDesc), C.sizeof(abi.RtsPulse),
              C.sizeof(abi.RtsBin), C.sizeof(abi.RtsStats), C.sizeof(lib.RtsPose), C.sizeof(abi.RtsResponse), C.sizeof(lib.RtsSizes)]
    assert list(sizes) == expect


def test_ray_record_layout():
    """PerRayData contract of /root/reference/ray_tracer.h:13-28: size 144, documented offsets."""
    dt = abi.RAY_RECORD
    assert dt.itemsize == 144
    offs = {n: dt.fields[n][1] for n in dt.names}
    assert offs == {"rayLength": 0, "refrIndex": 16, "reflDepth": 32, "refrDepth": 36, "maxRayIndex": 40, "rayDirection": 48,
                    "firstHitPoint": 72, "prevHitPoint": 96, "power": 120, "doppler": 128, "received": 136, "end": 140}


def test_result_sizes():
    """ray_tracer.cpp:600-626, 655: M = maxRefl+3 with refraction (forced to 2), else 1; D = maxRefl + maxRefr."""
    s = lib.result_sizes(abi.PulseSpec(grid=(4, 4, 4), max_refl=3, max_refr=0))
    assert (s.rays, s.ray_total, s.depth_total, s.slots, s.tri_cols) == (64, 64, 3, 1, 6)
    s = lib.result_sizes(abi.PulseSpec(grid=(1, 8, 8), max_refl=3, max_refr=7))
    assert (s.rays, s.ray_total, s.depth_total, s.slots) == (64, 64 * 6, 5, 6)
    p = abi.PulseSpec(grid=(1, 8, 8), max_refl=3, max_refr=1)
    assert (p.ray_total, p.depth_total, p.slots, p.tri_cols) == (64 * 6, 5, 6, 6)


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback():
    with pytest.raises(lib.RtsError, match="no CPU fallback|sm_100a"):
        lib.Engine(0)
