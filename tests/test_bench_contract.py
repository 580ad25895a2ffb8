"""bench.py's JSON contract: the reference arm runs on the CPU (the oracle port is its only engine), our arm refuses to run without a
GPU, and the committed bench line of the round carries every key the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from tests.util import ROOT

BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config")


def _bench(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "32")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in BASE_KEYS + ("impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_our_arm_needs_a_gpu():
    from aircraftoptimalcontrol_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    r = _bench("--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e", "--no-roofline", "--instances", "64", timeout=300)
    assert r.returncode != 0 and not any(ln.startswith("{") for ln in r.stdout.splitlines())   # no number without the CUDA path
    assert "no CUDA device" in r.stderr or "no CPU fallback" in r.stderr or "AcocError" in r.stderr


def test_committed_bench_line_has_every_contract_key():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench.json")))
    for k in BASE_KEYS + ("clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "trajectory_newton_iterations_per_second" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert d["e2e_host_refs"]["h2d_bytes_per_step"] > 1e6 and d["roofline"]["traffic"] > 0 and d["roofline"]["frac"] <= 1.2
    assert d["cpu_baseline"]["python_reference"]["value"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_under_torchrun_prints_one_line():
    """The driver launches the reference arm exactly like ours (torchrun for N > 1): rank 0 alone works and prints, the other ranks
    exit 0; the line carries OUR arm's config (the bounded CPU sample is described beside it)."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "32"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["gpu_launches"] == 0
    assert d["config"]["instances_per_gpu"] == 65536 and d["config"]["instances_total"] == 2 * 65536
    assert d["sample"]["instances"] == 32 and d["cpu_baseline"]["sample"].startswith("32 instances")
