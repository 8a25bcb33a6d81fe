"""The peer-memory bin exchange between processes (rts_comm_* over CUDA IPC, rts_b200.dist.PeerExchange) on every visible
GPU of the box — at least two — against the NCCL all-reduce pair and a single un-sharded engine; and its degenerate
single-rank form in this process."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from rts_b200 import lib as L, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_single_rank_exchange_is_the_identity(engine):
    """world = 1: publish + reduce over the one block leaves the accumulators as they were, and finalises them."""
    targets, spec = scenes.trihedral(n=160)
    engine.set_targets(targets)
    engine.trace(spec, L.RTS_OUT_BINS)
    want = engine.bins().copy()
    engine.comm_create(0, 1, 1 << 12)
    try:
        for _ in range(3):          # both halves of the block
            engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_FINALISE | L.RTS_ASYNC)
            engine.comm_allreduce_bins()
            got = engine.bins()
            # exact fields exact; the sums were accumulated with fp64 atomics in another order: last bits may differ
            assert all(np.array_equal(got[f], want[f]) for f in ("rx", "path", "npath", "min_slot", "own_min_slot", "direct"))
            assert all(np.allclose(got[f], want[f], rtol=1e-12, atol=0) for f in ("sum_sqrt_power", "sum_delay", "sum_phase", "sum_doppler", "power", "delay", "phase", "doppler"))
        with pytest.raises(L.RtsError):      # a table larger than the block was created for
            engine.comm_create(0, 1, 4)
            engine.trace(spec, L.RTS_OUT_BINS | L.RTS_NO_FINALISE)
            engine.comm_allreduce_bins()
    finally:
        engine.comm_destroy()
    with pytest.raises(L.RtsError):
        engine.comm_allreduce_bins()         # not set up


def test_peer_exchange_between_processes():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    n = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [json.loads(l.split("PEER_WORKER ", 1)[1]) for l in r.stdout.splitlines() if l.startswith("PEER_WORKER ")]
    assert r.returncode == 0 and len(lines) == n, r.stdout[-2000:] + r.stderr[-2000:]
    assert all(d["ok"] and d["pulses"] == 6 for d in lines), lines
    print("exchange device ms (idle GPUs, rank 0):", {k: lines[0].get(k) for k in ("peer_ms", "nccl_ms")})
