"""examples/rts_multigpu.cpp — a C++ host over the C-ABI with the bins reduced by NCCL (no Python in the loop): builds
against include/rts_b200.h + librts_b200.so, and on a GPU box the N-GPU reduced bins equal a single engine's."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")
BIN = os.path.join(EX, "rts_multigpu")

needs_nccl = pytest.mark.skipif(not os.path.exists("/usr/include/nccl.h") or shutil.which("g++") is None,
                                reason="system NCCL headers or g++ missing")


def _build():
    r = subprocess.run(["make", "-C", EX], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(BIN)


@needs_nccl
def test_cpp_host_builds_against_the_c_abi():
    _build()
    out = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "librts_b200.so" in out and "libnccl" in out


@needs_nccl
@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_cpp_host_reduced_bins_equal_single_engine(exchange):
    """All visible GPUs (one thread each); bins reduced by the library's peer-memory kernels (rts_comm_*) or by NCCL."""
    _build()
    r = subprocess.run([BIN, "--grid", "768", "--pulses", "3", "--cells", "96", "--exchange", exchange], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["exchange"] == exchange
    assert d["ok"] is True and d["bins"] >= 2 and d["max_rel_vs_single_gpu"] <= 1e-9 and d["kernel_launches_rank0"] > 0, d
