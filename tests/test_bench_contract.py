"""bench.py contract (task statement §④): the reference arm runs on the host cores without a GPU and prints one JSON
line with the agreed keys; the GPU arm's keys are checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-stride", "64"])
    assert COMMON <= set(out) and out["impl"] == "reference"
    assert out["unit"] == "Mrays/s" and out["value"] > 0 and out["higher_is_better"] is True and out["vs_baseline"] is None
    assert out["cpu_baseline"]["kind"] in ("port", "reference") and out["cpu_baseline"]["cores"] >= 1
    assert out["cpu_baseline"]["value"] == out["value"] == out["e2e"]["value"]
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in out["config"]


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    out = _run(["--steps", "4", "--warmup", "3", "--cpu-stride", "64", "--no-ncu"])   # the ncu child is exercised by tools/profile_round.sh
    assert COMMON | {"roofline", "clocks", "gpu_launches"} <= set(out)
    assert out["n_gpus"] == 1 and out["steps"] == 4 and out["warmup"] == 3 and out["scaling"] == "weak" and out["dtype"] == "f64"
    roof = out["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "algorithmic", "measured"} <= set(roof) and roof["unit"] == "GB/s"
    assert roof["bound"] in ("issue/latency", "hbm")
    # achieved / frac / traffic are measured DRAM figures of the dominant kernel (ncu counters); the byte model of
    # SURVEY.md §8(d) sits under "algorithmic".  A physically meaningful fraction never exceeds 1.
    if roof["frac"] is not None:
        assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-3 and 0 < roof["frac"] <= 1.05 and roof["traffic"] > 0
        assert 0 < roof["issue"]["frac"] <= 1.0
    alg = roof["algorithmic"]
    assert alg["bytes_per_segment"] == 1632.0 and alg["bytes_per_launch"] > 0 and alg["gbs"] > 0
    assert out["sustained"]["seconds"] >= 1.5 and out["sustained"]["value"] > 0
    assert out["gpu_launches"] > 0 and out["value"] > 0 and out["e2e"]["value"] > 0
    assert out["e2e"]["h2d_bytes_per_step"] > 0 and out["e2e"]["d2h_bytes_per_step"] > 0
    assert out["cpu_baseline"]["kind"] == "port" and "sample" in out["cpu_baseline"]
    assert "workload" in out["config"] and "model" not in out["config"]


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 without work or output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
