"""bench.py's output contract (one JSON line with the keys the driver reads) on small workloads, for every layout and
launch structure it can take - so that a code path of the benchmark that the default run does not visit cannot rot."""
import json
import os
import subprocess
import sys

import pytest

import common

BENCH = os.path.join(common.ROOT, "bench.py")
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline")


def run(*flags):
    r = subprocess.run([sys.executable, BENCH, *flags], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_contract():
    """--impl reference needs no GPU: the reference's CPU arithmetic on the host cores, same line shape."""
    d = run("--impl", "reference", "--molecules", "3000", "--steps", "3", "--warmup", "1")
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "body-steps/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [
    (),
    ("--no-fuse",),
    ("--graph",),
    ("--layout", "soa", "--shuffle"),
    ("--layout", "openmm-mixed", "--shuffle"),
    ("--layout", "openmm-double", "--shuffle", "atoms"),
    ("--mode", "10"),
    ("--workload", "mixed"),
    ("--workload", "mixed", "--layout", "openmm-mixed", "--graph"),
    ("--forces", "constant", "--dt-fs", "2"),
])
def test_bench_line_contract_small_workloads(flags):
    lean = () if flags == () else ("--no-cpu-baseline", "--no-gpu-reference")     # the CPU / reference-CUDA legs once, in the default case
    d = run("--molecules", "20000", "--steps", "12", "--warmup", "3", *flags, *lean)
    for key in REQUIRED:
        assert key in d, key
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0 and d["n_gpus"] == 1
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and 0.0 < r["frac"] < 1.0                  # a fraction of a bandwidth
    assert "workload" in d["config"] and "model" not in d["config"]
    p = d["parity_subsample"]
    assert p["ok"] and p["p99_rel_R"] <= 1e-6 and p["p99_rel_V"] <= 1e-6
    if d["e2e"] is not None:
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    if d["cpu_baseline"] is not None:
        assert d["cpu_baseline"]["value"] > 0
