"""The measurement contract of bench.py, checked without a GPU: the committed bench lines carry every key the driver reads,
the reference arm runs on the host cores (bounded sample) and prints the same shape, and non-zero ranks of the reference
arm exit without work."""
import json
import os
import subprocess
import sys

import pytest

from tests.helpers import ROOT

OURS_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]


@pytest.mark.parametrize("name,n", [("bench_r1_n1.json", 1), ("bench_r1_n2.json", 2), ("bench_r1_n4.json", 4), ("bench_r1_n8.json", 8)])
def test_committed_bench_lines_follow_the_contract(name, n):
    line = json.load(open(os.path.join(ROOT, "profiles", name)))
    for k in OURS_KEYS:
        assert k in line, k
    assert line["n_gpus"] == n and line["unit"] == "Msamples/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["vs_baseline"] is None                  # BASELINE.md publishes no number for this metric
    assert line["dtype"] == "f32" and "workload" in line["config"] and "model" not in line["config"]
    assert line["warmup"] >= 3 and line["steps"] >= 1 and line["value"] > 0 and line["gpu_launches"] > 0
    e = line["e2e"]
    assert e["unit"] == line["unit"] and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != line["value"]                  # measured separately, host buffers inside the timed region
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = line["clocks"]
    assert c["sm_mhz"] > 0 and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        b = line["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["unit"] == line["unit"] and b["sample"]
        assert line["value"] / b["value"] > 100       # sanity: the GPU path is not the CPU path in disguise


def test_reference_arm_prints_one_line_and_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] == line["value"]
